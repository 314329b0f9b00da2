"""jumanji==0.2.2 jumanji/environments/routing/connector/utils.py (UPSTREAM, restated from the
published source; the reference imports these six helpers at parallel_random_walk.py:37-44 and
mirrors is_valid_position at :401-429)."""
import jax
import jax.numpy as jnp

from .constants import AGENT_INITIAL_VALUE, DOWN, EMPTY, LEFT, NOOP, PATH, POSITION, RIGHT, TARGET, UP  # noqa: F401
from .types import Agent


def get_path(agent_id):
    return 1 + 3 * agent_id


def get_position(agent_id):
    return 2 + 3 * agent_id


def get_target(agent_id):
    return 3 + 3 * agent_id


def is_target(value):
    return (value > 0) & ((value - TARGET) % 3 == 0)


def is_position(value):
    return (value > 0) & ((value - POSITION) % 3 == 0)


def is_path(value):
    return (value > 0) & ((value - PATH) % 3 == 0)


def get_agent_id(value):
    return 0 if value == 0 else (value - 1) // 3


def move_position(position, action):
    row, col = position[0], position[1]
    move_noop = lambda row, col: jnp.array([row, col], jnp.int32)  # noqa: E731
    move_left = lambda row, col: jnp.array([row, col - 1], jnp.int32)  # noqa: E731
    move_up = lambda row, col: jnp.array([row - 1, col], jnp.int32)  # noqa: E731
    move_right = lambda row, col: jnp.array([row, col + 1], jnp.int32)  # noqa: E731
    move_down = lambda row, col: jnp.array([row + 1, col], jnp.int32)  # noqa: E731
    return jax.lax.switch(action, [move_noop, move_up, move_right, move_down, move_left], row, col)


def move_agent(agent, grid, new_pos):
    grid = grid.at[tuple(agent.position)].set(get_path(agent.id))
    grid = grid.at[tuple(new_pos)].set(get_position(agent.id))
    new_agent = Agent(id=agent.id, start=agent.start, target=agent.target, position=jnp.array(new_pos))
    return new_agent, grid


def is_valid_position(grid, agent, position):
    row, col = position[0], position[1]
    grid_size = grid.shape[0]
    in_bounds = (0 <= row) & (row < grid_size) & (0 <= col) & (col < grid_size)
    open_cell = (grid[row, col] == EMPTY) | (grid[row, col] == get_target(agent.id))
    not_connected = ~agent.connected
    return in_bounds & open_cell & not_connected


def connected_or_blocked(agent, action_mask):
    return agent.connected.all() | jnp.logical_not(action_mask[1:].any())


def get_agent_grid(agent_id, grid):
    position = get_position(agent_id)
    target = get_target(agent_id)
    path = get_path(agent_id)
    agent_head = (grid == position) * position
    agent_target = (grid == target) * target
    agent_path = (grid == path) * path
    return agent_head + agent_target + agent_path


def get_correction_mask(old_grid, joined_grid, agent_id):
    """An agent whose head is missing from the joined grid has collided; adding the mask puts the
    head back on its old cell (which the join holds as PATH = POSITION - 1)."""
    position = get_position(agent_id)
    agent_collided = ~jnp.any(joined_grid == position)
    correction_mask = jnp.where(agent_collided, (old_grid == position) * 1, jnp.zeros_like(old_grid))
    return correction_mask, agent_collided


def switch_perspective(grid, agent_id, num_agents):
    """Encodes the observation with respect to the current agent defined by `agent_id`: in each agent's
    observation its own path / position / target are 1, 2, 3 and the other agents follow cyclically."""
    new_grid = grid - AGENT_INITIAL_VALUE  # Center on first agent
    new_grid -= 3 * agent_id  # Center on current agent
    new_grid %= 3 * num_agents  # Values of other agents wrap around
    new_grid += AGENT_INITIAL_VALUE
    # Take care of zeros: empty cells stay empty
    return jnp.where(grid == EMPTY, EMPTY, new_grid)
