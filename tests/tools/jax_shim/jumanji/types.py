"""jumanji==0.2.2 jumanji/types.py (UPSTREAM, restated from the published source): StepType,
TimeStep and the restart / transition / termination / truncation constructors."""
from typing import Any, Dict, Optional, Sequence, Union

import chex
import jax.numpy as jnp


class StepType:
    """Defines the status of a `TimeStep` within a sequence (int8 scalars upstream)."""

    FIRST = jnp.array(0, jnp.int8)  # the first `TimeStep` in a sequence
    MID = jnp.array(1, jnp.int8)    # any `TimeStep` that is not FIRST or LAST
    LAST = jnp.array(2, jnp.int8)   # the last `TimeStep` in a sequence


@chex.dataclass
class TimeStep:
    step_type: Any
    reward: Any
    discount: Any
    observation: Any
    extras: Optional[Dict] = None

    def first(self):
        return self.step_type == StepType.FIRST

    def mid(self):
        return self.step_type == StepType.MID

    def last(self):
        return self.step_type == StepType.LAST


def restart(observation, extras: Optional[Dict] = None, shape: Union[int, Sequence[int]] = ()) -> TimeStep:
    """`TimeStep` with `step_type` FIRST: zero reward, discount one."""
    extras = extras or {}
    return TimeStep(step_type=StepType.FIRST, reward=jnp.zeros(shape, dtype=float), discount=jnp.ones(shape, dtype=float), observation=observation, extras=extras)


def transition(reward, observation, discount=None, extras: Optional[Dict] = None, shape: Union[int, Sequence[int]] = ()) -> TimeStep:
    """`TimeStep` with `step_type` MID; discount defaults to one."""
    discount = discount if discount is not None else jnp.ones(shape, dtype=float)
    extras = extras or {}
    return TimeStep(step_type=StepType.MID, reward=reward, discount=discount, observation=observation, extras=extras)


def termination(reward, observation, extras: Optional[Dict] = None, shape: Union[int, Sequence[int]] = ()) -> TimeStep:
    """`TimeStep` with `step_type` LAST and discount zero."""
    extras = extras or {}
    return TimeStep(step_type=StepType.LAST, reward=reward, discount=jnp.zeros(shape, dtype=float), observation=observation, extras=extras)


def truncation(reward, observation, discount=None, extras: Optional[Dict] = None, shape: Union[int, Sequence[int]] = ()) -> TimeStep:
    """`TimeStep` with `step_type` LAST that keeps the discount."""
    discount = discount if discount is not None else jnp.ones(shape, dtype=float)
    extras = extras or {}
    return TimeStep(step_type=StepType.LAST, reward=reward, discount=discount, observation=observation, extras=extras)
