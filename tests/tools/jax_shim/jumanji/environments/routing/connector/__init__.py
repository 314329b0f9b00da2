from .types import Agent, State, Observation  # noqa: F401
