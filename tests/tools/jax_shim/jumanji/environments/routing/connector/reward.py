"""jumanji==0.2.2 jumanji/environments/routing/connector/reward.py (UPSTREAM, restated from the published
source): DenseRewardFn."""
import abc

import jax.numpy as jnp


class RewardFn(abc.ABC):
    @abc.abstractmethod
    def __call__(self, state, action, next_state):
        """The reward function used in the `Connector` environment."""


class DenseRewardFn(RewardFn):
    """Returns: reward of 1.0 * `connected_reward` for each agent that connects on that step, and adds
    `timestep_reward` for each agent that has not connected yet (upstream defaults via the env: -0.03, 0.1)."""

    def __init__(self, timestep_reward: float = -0.03, connected_reward: float = 0.1) -> None:
        self.timestep_reward = timestep_reward
        self.connected_reward = connected_reward

    def __call__(self, state, action, next_state):
        connected_rewards = self.connected_reward * jnp.asarray(~state.agents.connected & next_state.agents.connected, float)
        timestep_rewards = self.timestep_reward * jnp.asarray(~state.agents.connected, float)
        return connected_rewards + timestep_rewards
