"""jumanji==0.2.2 connector types.  `Agent` is the reference's own mirror of the upstream class
(board_generation/types_addition.py:16-42: "This file mirrors the agent class from types.py")."""
import chex
from routing_board_generation.board_generation_methods.jax_implementation.board_generation.types_addition import Agent  # noqa: F401


@chex.dataclass
class State:
    """field order as printed in package_evaluation/profiling_generators.ipynb cell 4"""

    grid: chex.Array
    step_count: chex.Array
    agents: Agent
    key: chex.PRNGKey


@chex.dataclass
class Observation:
    grid: chex.Array
    action_mask: chex.Array
    step_count: chex.Array
