"""Connector environment + wrappers on the CUDA engine.

Mirrors jumanji==0.2.2 (UPSTREAM, requirements.txt:5) as the reference uses it:
  Connector(generator, reward_fn, time_limit)       rl_training/setup_train.py:158
  reset(key) -> (State, TimeStep), step(state, action) -> (State, TimeStep)
  _get_action_mask / _obs_from_grid / _get_extras   demos/board_generator_demo.py:83-96
  MultiToSingleWrapper, VmapAutoResetWrapper        rl_training/setup_train.py:160,166
Batching: the leading axis of keys / state leaves is the vmap axis; a single
key of shape (2,) gives single-env (unbatched) pytrees like the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import engine
from .online_generators import Generator, UniformRandomGenerator
from .types import Agent, Observation, State, TimeStep


@dataclass
class DenseRewardFn:
    """jumanji DenseRewardFn: connected_reward when an agent connects, timestep_reward per unconnected step."""

    timestep_reward: float = -0.03
    connected_reward: float = 0.1


def _batch(tree, batched: bool):
    return tree if batched else tree.map(lambda t: t[None])


def _unbatch(tree, batched: bool):
    return tree if batched else tree[0]


class Connector:
    def __init__(self, generator: Optional[Generator] = None, reward_fn: Optional[DenseRewardFn] = None, time_limit: int = 50, viewer=None) -> None:
        self._generator = generator or UniformRandomGenerator(grid_size=10, num_agents=5)
        if not getattr(self._generator, "kind", ""):
            raise TypeError("generator must be one of this package's kernel-backed generators (Uniform / ParallelRandomWalk / SeedExtension)")
        self._reward_fn = reward_fn or DenseRewardFn()
        self.time_limit = time_limit
        self.num_agents = self._generator.num_agents
        self.grid_size = self._generator.grid_size
        self._agent_ids = None
        self._viewer = viewer

    def __repr__(self) -> str:
        return f"Connector(grid_size={self.grid_size}, num_agents={self.num_agents}, time_limit={self.time_limit})"

    @property
    def unwrapped(self) -> "Connector":
        return self

    # -- public API ---------------------------------------------------------
    def reset(self, key) -> Tuple[State, TimeStep]:
        keys, batched = engine.as_keys(key)
        st, ts = engine.connector_reset(self._generator.kind, keys, self.grid_size, self.num_agents)
        return _unbatch(st, batched), _unbatch(ts, batched)

    def step(self, state: State, action) -> Tuple[State, TimeStep]:
        batched = state.grid.dim() == 3
        st = _batch(state, batched)
        new, ts = engine.connector_step(st, action, self.time_limit, self._reward_fn.timestep_reward, self._reward_fn.connected_reward)
        return _unbatch(new, batched), _unbatch(ts, batched)

    # -- the private trio the reference's own scripts call ------------------
    def _observe(self, state: State) -> TimeStep:
        batched = state.grid.dim() == 3
        return _unbatch(engine.connector_observe(_batch(state, batched)), batched)

    def _get_action_mask_all(self, agents: Agent, grid: torch.Tensor) -> torch.Tensor:
        """vmap(_get_action_mask, (0, None))(agents, grid) -> bool[N,5] (or [B,N,5])."""
        batched = grid.dim() == 3
        st = self._state_from(agents, grid, batched)
        return _unbatch(engine.connector_observe(st), batched).observation.action_mask

    def _get_action_mask(self, agent: Agent, grid: torch.Tensor) -> torch.Tensor:
        """One agent's mask bool[5] on `grid` (jumanji env.py _get_action_mask)."""
        aid = int(agent.id)
        n = self.num_agents
        dev = grid.device
        pad = lambda v, fill: torch.full((n, 2), fill, dtype=torch.int32, device=dev).index_put((torch.tensor(aid, device=dev),), v.to(torch.int32))
        agents = Agent(id=torch.arange(n, dtype=torch.int32, device=dev), start=pad(agent.start, -1), target=pad(agent.target, -2), position=pad(agent.position, 0))
        return self._get_action_mask_all(agents, grid)[aid]

    def _obs_from_grid(self, grid: torch.Tensor) -> torch.Tensor:
        batched = grid.dim() == 3
        n = self.num_agents
        g = grid if batched else grid[None]
        z = torch.zeros((g.shape[0], n, 2), dtype=torch.int32, device=g.device)
        agents = Agent(id=torch.arange(n, dtype=torch.int32, device=g.device).expand(g.shape[0], n).contiguous(), start=z, target=z - 1, position=z)
        ts = engine.connector_observe(self._state_from(agents, g, True))
        return ts.observation.grid if batched else ts.observation.grid[0]

    def _get_extras(self, state: State) -> dict:
        return self._observe(state).extras

    def _state_from(self, agents: Agent, grid: torch.Tensor, batched: bool) -> State:
        g = grid if batched else grid[None]
        ag = agents if batched else agents.map(lambda t: t[None])
        B = g.shape[0]
        return State(key=torch.zeros((B, 2), dtype=torch.uint32, device=g.device), grid=g.to(torch.int32).contiguous(), step_count=torch.zeros((B,), dtype=torch.int32, device=g.device), agents=ag.map(lambda t: t.to(torch.int32).contiguous()))


class VmapAutoResetWrapper:
    """jumanji.wrappers.VmapAutoResetWrapper: on LAST, `key, _ = split(state.key)`, reset(key), keep the
    terminal reward / discount / step_type / extras and swap in the reset observation.
    The stepping, the compaction of finished envs and their regeneration run as two launches."""

    def __init__(self, env: Connector):
        self._env = env

    @property
    def unwrapped(self) -> Connector:
        return self._env.unwrapped

    def __getattr__(self, name):
        return getattr(self._env, name)

    def reset(self, key):
        keys, _ = engine.as_keys(key)
        return self._env.reset(keys)

    def step(self, state: State, action, inplace: bool = False):
        e = self._env.unwrapped
        return engine.connector_step(state, action, e.time_limit, e._reward_fn.timestep_reward, e._reward_fn.connected_reward, autoreset_kind=e._generator.kind, inplace=inplace, owner=e)

    def step_random(self, state: State, inplace: bool = False):
        """Random-policy step in the same launch (the agent=random benchmark loop)."""
        e = self._env.unwrapped
        return engine.connector_step(state, None, e.time_limit, e._reward_fn.timestep_reward, e._reward_fn.connected_reward, autoreset_kind=e._generator.kind, inplace=inplace, random_policy=True, owner=e)


    def rollout_random(self, state: State, n_steps: int, out: Optional[TimeStep] = None):
        """`n_steps` random-policy steps with auto-reset in one library call: generation, reset and
        stepping fused into one launch sequence.  Updates `state` in place; returns
        (state, TimeStep stacked [n_steps, B, ...], actions[n_steps, B, N])."""
        e = self._env.unwrapped
        return engine.connector_rollout_random(state, n_steps, e.time_limit, e._reward_fn.timestep_reward, e._reward_fn.connected_reward, autoreset_kind=e._generator.kind, out=out, owner=e)


class MultiToSingleWrapper:
    """jumanji.wrappers.MultiToSingleWrapper: reward -> sum over agents, discount -> max over agents."""

    def __init__(self, env, reward_aggregator=torch.sum, discount_aggregator=torch.amax):
        self._env = env
        self._ragg = reward_aggregator
        self._dagg = discount_aggregator

    @property
    def unwrapped(self):
        return self._env.unwrapped

    def __getattr__(self, name):
        return getattr(self._env, name)

    def _aggregate(self, ts: TimeStep) -> TimeStep:
        return ts.replace(reward=self._ragg(ts.reward, dim=-1), discount=self._dagg(ts.discount, dim=-1))

    def reset(self, key):
        st, ts = self._env.reset(key)
        return st, self._aggregate(ts)

    def step(self, state, action, **kw):
        st, ts = self._env.step(state, action, **kw)
        return st, self._aggregate(ts)


def make_random_policy_connector():
    """Uniform over the legal actions (NOOP included), as jumanji's make_random_policy_connector
    (rl_training/setup_train.py:246) in distribution; see include/rbg_b200.h rbg_random_actions."""

    def policy(state: State) -> torch.Tensor:
        batched = state.grid.dim() == 3
        act = engine.random_actions(_batch(state, batched))
        return act if batched else act[0]

    return policy
