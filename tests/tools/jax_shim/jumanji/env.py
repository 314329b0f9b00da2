"""jumanji==0.2.2 jumanji/env.py (UPSTREAM, restated): the Environment / Wrapper base classes as far as the
Connector call sites of the reference use them (rl_training/setup_train.py:158-166,400)."""
import abc


class Environment(abc.ABC):
    def __repr__(self) -> str:
        return "Environment."

    @abc.abstractmethod
    def reset(self, key):
        """-> (state, timestep)"""

    @abc.abstractmethod
    def step(self, state, action):
        """-> (state, timestep)"""

    @property
    def unwrapped(self) -> "Environment":
        return self


class Wrapper(Environment):
    """Wraps the environment to allow modular transformations (upstream jumanji/wrappers.py)."""

    def __init__(self, env: Environment):
        super().__init__()
        self._env = env

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({repr(self._env)})"

    def __getattr__(self, name: str):
        if name == "__setstate__":
            raise AttributeError(name)
        return getattr(self._env, name)

    @property
    def unwrapped(self) -> Environment:
        return self._env.unwrapped

    def reset(self, key):
        return self._env.reset(key)

    def step(self, state, action):
        return self._env.step(state, action)
