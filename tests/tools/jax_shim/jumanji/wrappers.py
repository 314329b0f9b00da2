"""jumanji==0.2.2 jumanji/wrappers.py (UPSTREAM, restated from the published source), the two wrappers the
reference stacks on Connector: `VmapAutoResetWrapper(MultiToSingleWrapper(Connector(generator)))`
(rl_training/setup_train.py:158-166)."""
from typing import Callable, Tuple

import jax
import jax.numpy as jnp

from jumanji.env import Environment, Wrapper
from jumanji.types import TimeStep


class MultiToSingleWrapper(Wrapper):
    """A wrapper that converts a multi-agent Environment to a single-agent Environment."""

    def __init__(self, env: Environment, reward_aggregator: Callable = jnp.sum, discount_aggregator: Callable = jnp.max):
        super().__init__(env)
        self._reward_aggregator = reward_aggregator
        self._discount_aggregator = discount_aggregator

    def _aggregate_timestep(self, timestep: TimeStep) -> TimeStep:
        """Apply the reward and discount aggregator to a multi-agent timestep to create a new timestep
        that consists of a scalar reward and discount value."""
        return TimeStep(
            step_type=timestep.step_type,
            observation=timestep.observation,
            reward=self._reward_aggregator(timestep.reward),
            discount=self._discount_aggregator(timestep.discount),
            extras=timestep.extras,
        )

    def reset(self, key):
        state, timestep = self._env.reset(key)
        timestep = self._aggregate_timestep(timestep)
        return state, timestep

    def step(self, state, action):
        state, timestep = self._env.step(state, action)
        timestep = self._aggregate_timestep(timestep)
        return state, timestep


class VmapAutoResetWrapper(Wrapper):
    """Efficient combination of VmapWrapper and AutoResetWrapper: resets ONLY the terminated environments.
    Takes batched `keys` / `state` / `action`; the State must carry a `key`."""

    def reset(self, key):
        state, timestep = jax.vmap(self._env.reset)(key)
        return state, timestep

    def step(self, state, action):
        # Vmap homogeneous computation (parallelizable).
        state, timestep = jax.vmap(self._env.step)(state, action)
        # Map heterogeneous computation (non-parallelizable).
        state, timestep = jax.lax.map(lambda args: self._maybe_reset(*args), (state, timestep))
        return state, timestep

    def _auto_reset(self, state, timestep: TimeStep) -> Tuple:
        """Reset the state and overwrite `timestep.observation` with the reset observation."""
        if not hasattr(state, "key"):
            raise AttributeError("This wrapper assumes that the state has attribute key which is used as the source of randomness for automatic reset")
        # Make sure that the random key in the environment changes at each call to reset.
        # State is a type variable hence it does not have key type hinted, so we type ignore.
        key, _ = jax.random.split(state.key)
        state, reset_timestep = self._env.reset(key)
        # Replace observation with reset observation.
        timestep = timestep.replace(observation=reset_timestep.observation)
        return state, timestep

    def _maybe_reset(self, state, timestep: TimeStep) -> Tuple:
        """Overwrite the state and timestep appropriately if the episode terminates."""
        state, timestep = jax.lax.cond(timestep.last(), self._auto_reset, lambda st, ts: (st, ts), state, timestep)
        return state, timestep
