// connector_device.cuh -- Connector observation / action-mask helpers shared
// by the step kernel and the generator kernel's fused auto-reset epilogue.
//
// Semantics: jumanji==0.2.2 Connector (UPSTREAM; reference call sites
// rl_training/setup_train.py:158-166, demos/board_generator_demo.py:83-96)
// and its mirror in the reference, parallel_random_walk.py:401-429.
#pragma once

#include "rbg_device.cuh"

namespace rbg {

// grid held in shared memory as uint8, G x G, row stride S, origin offset so
// that (r,c) = base[(r+pad)*S + c+pad].  pad=2 boards carry a 0xFF border
// (never EMPTY, never anybody's code) so neighbour reads need no bounds test.
struct SmemGrid {
  const uint8_t *base;
  int S;
  int pad;
  int G;
  __device__ __forceinline__ uint32_t at(int r, int c) const {
    return base[(r + pad) * S + (c + pad)];
  }
};

// is_valid_position for the four moves in action order UP, RIGHT, DOWN, LEFT.
// bit (a-1) set <=> action a legal.  `connected` agents have no legal move.
__device__ __forceinline__ uint32_t move_mask(const SmemGrid &g, int r, int c,
                                              int agent, bool connected) {
  const uint32_t tgt = 3u * agent + TARGET;
  // (r, c) is on the grid; a neighbour off the grid reads as 0xFF (never empty, nobody's target)
  const uint8_t *pc = g.base + (r + g.pad) * g.S + (c + g.pad);
  const uint32_t u = r > 0 ? pc[-g.S] : 0xFFu, ri = c < g.G - 1 ? pc[1] : 0xFFu;
  const uint32_t d = r < g.G - 1 ? pc[g.S] : 0xFFu, l = c > 0 ? pc[-1] : 0xFFu;
  uint32_t m = 0;
  m |= (u == 0u || u == tgt) ? 1u : 0u;
  m |= (ri == 0u || ri == tgt) ? 2u : 0u;
  m |= (d == 0u || d == tgt) ? 4u : 0u;
  m |= (l == 0u || l == tgt) ? 8u : 0u;
  return connected ? 0u : m;
}

// action_mask[5] bytes of one agent: [1, UP, RIGHT, DOWN, LEFT]
__device__ __forceinline__ void store_mask5(uint8_t *dst, uint32_t m) {
  dst[0] = 1;
  dst[1] = (uint8_t)(m & 1u);
  dst[2] = (uint8_t)((m >> 1) & 1u);
  dst[3] = (uint8_t)((m >> 2) & 1u);
  dst[4] = (uint8_t)((m >> 3) & 1u);
}

// Write one env's observation.grid [N,G,G] from a shared-memory grid with a
// warp: 128-bit stores when cells % 4 == 0 (then every 16-byte chunk lies in
// one agent slice and dst is 16-byte aligned), scalar stores otherwise.
__device__ __forceinline__ void warp_write_obs(const SmemGrid &g, int N,
                                               const FastDiv &divG,
                                               int32_t *__restrict__ dst,
                                               int lane) {
  const int G = g.G, cells = G * G, n3 = 3 * N;
  if ((cells & 3) == 0) {
    const int c4 = cells >> 2;
    for (int a = 0; a < N; ++a) {
      int4 *o = reinterpret_cast<int4 *>(dst + (size_t)a * cells);
      for (int q = lane; q < c4; q += 32) {
        uint32_t r, c;
        divG.divmod((uint32_t)(4 * q), r, c);
        int v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = obs_value((int)g.at((int)r, (int)c), 3 * a, n3);
          if (++c == (uint32_t)G) {
            c = 0;
            ++r;
          }
        }
        o[q] = make_int4(v[0], v[1], v[2], v[3]);
      }
    }
  } else {
    for (int i = lane; i < N * cells; i += 32) {
      const int a = i / cells, cell = i - a * cells;
      uint32_t r, c;
      divG.divmod((uint32_t)cell, r, c);
      dst[i] = obs_value((int)g.at((int)r, (int)c), 3 * a, n3);
    }
  }
}

}  // namespace rbg
