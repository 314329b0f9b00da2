"""jumanji==0.2.2 jumanji/environments/routing/connector/env.py (UPSTREAM, not under /root/reference;
restated from the published source, written against the jax API exactly like upstream and executed on the
NumPy stand-in).  An independent restatement: nothing here was derived from oracle/rbg_oracle.c or from
the CUDA kernels; agreement between the three is N-version agreement, not an upstream pin.

The agent-stepping rule is additionally mirrored by the reference itself
(parallel_random_walk.py:101-145 `_step_agents`, :376-429 `_step_agent` / `_is_valid_position`,
"mirrors ... Connector"), and the reset recipe by demos/board_generator_demo.py:80-97.
"""
from typing import Dict, Optional, Tuple

import chex
import jax
import jax.numpy as jnp

from jumanji.env import Environment
from jumanji.environments.routing.connector.constants import AGENT_INITIAL_VALUE, NOOP, PATH
from jumanji.environments.routing.connector.reward import DenseRewardFn, RewardFn
from jumanji.environments.routing.connector.types import Agent, Observation, State
from jumanji.environments.routing.connector.utils import (
    connected_or_blocked,
    get_agent_grid,
    get_correction_mask,
    is_valid_position,
    move_agent,
    move_position,
    switch_perspective,
)
from jumanji.types import TimeStep, restart, termination, transition


class Connector(Environment):
    """The `Connector` environment: a multi-agent gridworld in which every agent must connect its start to
    its target without crossing another agent's path.

    - observation: grid int32 (num_agents, grid_size, grid_size), each slice the grid from one agent's
      perspective (its own path / position / target are 1, 2, 3); action_mask bool (num_agents, 5);
      step_count int32 ().
    - action: int32 (num_agents,): 0 no-op, 1 up, 2 right, 3 down, 4 left.
    - reward: float (num_agents,), dense: +connected_reward on connecting, timestep_reward per step until then.
    - episode termination: all agents connected or blocked, or `time_limit` steps.
    """

    def __init__(self, generator=None, reward_fn: Optional[RewardFn] = None, time_limit: int = 50, viewer=None) -> None:
        if generator is None:  # upstream default: UniformRandomGenerator(grid_size=10, num_agents=5); the reference carries a copy
            from routing_board_generation.rl_training.online_generators.uniform_generator import UniformRandomGenerator

            generator = UniformRandomGenerator(grid_size=10, num_agents=5)
        self._generator = generator
        self._reward_fn = reward_fn or DenseRewardFn(timestep_reward=-0.03, connected_reward=0.1)
        self.time_limit = time_limit
        self.num_agents = self._generator.num_agents
        self.grid_size = self._generator.grid_size
        self._agent_ids = jnp.arange(self.num_agents)
        self._viewer = viewer

    def reset(self, key: chex.PRNGKey) -> Tuple[State, TimeStep]:
        state = self._generator(key)
        action_mask = jax.vmap(self._get_action_mask, (0, None))(state.agents, state.grid)
        observation = Observation(grid=self._obs_from_grid(state.grid), action_mask=action_mask, step_count=state.step_count)
        extras = self._get_extras(state)
        timestep = restart(observation=observation, extras=extras, shape=(self.num_agents,))
        return state, timestep

    def step(self, state: State, action: chex.Array) -> Tuple[State, TimeStep]:
        agents, grid = self._step_agents(state, action)
        new_state = State(grid=grid, step_count=state.step_count + 1, agents=agents, key=state.key)

        # Construct timestep: get observations, rewards, discounts
        grids = self._obs_from_grid(grid)
        reward = self._reward_fn(state, action, new_state)
        action_mask = jax.vmap(self._get_action_mask, (0, None))(agents, grid)
        observation = Observation(grid=grids, action_mask=action_mask, step_count=new_state.step_count)

        dones = jax.vmap(connected_or_blocked)(agents, action_mask)
        discount = jnp.asarray(jnp.logical_not(dones), dtype=float)
        extras = self._get_extras(new_state)
        timestep = jax.lax.cond(
            dones.all() | (new_state.step_count >= self.time_limit),
            lambda: termination(reward=reward, observation=observation, extras=extras, shape=self.num_agents),
            lambda: transition(reward=reward, observation=observation, extras=extras, discount=discount, shape=self.num_agents),
        )
        return new_state, timestep

    def _step_agents(self, state: State, action: chex.Array) -> Tuple[Agent, chex.Array]:
        """Steps all agents at the same time correcting for possible collisions.

        If a collision occurs we place the agent with the lower `agent_id` in its previous position."""
        agent_ids = jnp.arange(self.num_agents)
        # Step all agents at the same time (separately) and return all of the grids
        agents, grids = jax.vmap(self._step_agent, in_axes=(0, None, 0))(state.agents, state.grid, action)

        # Get grids with only values related to a single agent.
        agent_grids = jax.vmap(get_agent_grid)(agent_ids, grids)
        joined_grid = jnp.max(agent_grids, 0)  # join the grids

        # Create a correction mask for possible collisions (see the docs of `get_correction_mask`)
        correction_fn = jax.vmap(get_correction_mask, in_axes=(None, None, 0))
        correction_masks, collided_agents = correction_fn(state.grid, joined_grid, agent_ids)
        correction_mask = jnp.sum(correction_masks, 0)

        # Correct state.agents: old agents where they collided, new agents otherwise
        agents = jax.vmap(lambda collided, old_agent, new_agent: jax.lax.cond(collided, lambda: old_agent, lambda: new_agent))(
            collided_agents, state.agents, agents
        )
        # Create the new grid by fixing old one with correction mask and adding the obstacles
        return agents, joined_grid + correction_mask

    def _step_agent(self, agent: Agent, grid: chex.Array, action) -> Tuple[Agent, chex.Array]:
        """Moves the agent according to the given action if it is possible."""
        new_pos = move_position(agent.position, action)
        new_agent, new_grid = jax.lax.cond(
            is_valid_position(grid, agent, new_pos) & (action != NOOP),
            move_agent,
            lambda *_: (agent, grid),
            agent,
            grid,
            new_pos,
        )
        return new_agent, new_grid

    def _obs_from_grid(self, grid: chex.Array) -> chex.Array:
        """Gets the observation vector for all agents."""
        return jax.vmap(switch_perspective, (None, 0, None))(grid, self._agent_ids, self.num_agents)

    def _get_action_mask(self, agent: Agent, grid: chex.Array) -> chex.Array:
        """Gets an agent's action mask."""
        # Don't check action 0 because no-op is always valid
        actions = jnp.arange(1, 5)

        def is_valid_action(action):
            agent_pos = move_position(agent.position, action)
            return is_valid_position(grid, agent, agent_pos)

        mask = jnp.ones(5, dtype=bool)
        mask = mask.at[actions].set(jax.vmap(is_valid_action)(actions))
        return mask

    def _get_extras(self, state: State) -> Dict:
        """Computes extras metrics to be returned within the timestep."""
        offset = AGENT_INITIAL_VALUE
        total_path_length = jnp.sum((offset + (state.grid - offset) % 3) == PATH)
        # Add agents' head
        total_path_length += self.num_agents
        extras = {
            "num_connections": jnp.sum(state.agents.connected),
            "ratio_connections": jnp.mean(state.agents.connected),
            "total_path_length": total_path_length,
        }
        return extras
