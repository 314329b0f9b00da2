// select.cuh -- the first `nsel` entries of jax.random._shuffle(key, arange(n)),
// warp-cooperative.  Shared by the ParallelRandomWalk / Uniform kernel (start
// cells, parallel_random_walk.py:160, uniform_generator.py:82) and the
// SeedExtension seeding kernel (seed_extension.py:131-135).
#pragma once

#include "rbg_device.cuh"

namespace rbg {

// ---- start-cell selection: the first `nsel` entries of
// jax.random._shuffle(key, arange(n)) = the nsel smallest (sort_key, index)
// pairs in order, sort_key = random_bits(sub, (n,)).

// exact, storage-free: nsel rounds of "smallest composite greater than the last"
__device__ inline void select_exact(uint32_t sk0, uint32_t sk1, int n, int nsel,
                             uint16_t *out, int lane) {
  const int h = (n + 1) >> 1;
  uint64_t last = 0;
  for (int rnk = 0; rnk < nsel; ++rnk) {
    uint64_t best = ~0ull;
    for (int j0 = 0; j0 < h; j0 += 32) {
      const int j = j0 + lane;
      if (j < h) {
        const bool has1 = (j + h) < n;
        uint32_t o0, o1;
        tf_block(sk0, sk1, (uint32_t)j, has1 ? (uint32_t)(j + h) : 0u, o0, o1);
        const uint64_t c0 = ((uint64_t)o0 << 32) | (uint32_t)j;
        if ((rnk == 0 || c0 > last) && c0 < best) best = c0;
        if (has1) {
          const uint64_t c1 = ((uint64_t)o1 << 32) | (uint32_t)(j + h);
          if ((rnk == 0 || c1 > last) && c1 < best) best = c1;
        }
      }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const uint64_t other = __shfl_xor_sync(FULL, best, off);
      best = other < best ? other : best;
    }
    if (lane == 0) out[rnk] = (uint16_t)(best & 0xffffu);
    last = best;
  }
  __syncwarp();
}

__device__ inline void select_smallest(uint32_t sk0, uint32_t sk1, int n, int nsel,
                                uint32_t thresh, int cap, uint64_t *cand,
                                uint16_t *out, int lane, bool force_exact) {
  const int h = (n + 1) >> 1;
  const uint32_t lt = (1u << lane) - 1u;
  int cnt = 0;
  for (int j0 = 0; j0 < h; j0 += 32) {
    const int j = j0 + lane;
    const bool act = j < h;
    const bool has1 = act && (j + h) < n;
    uint32_t o0, o1;
    tf_block(sk0, sk1, (uint32_t)j, has1 ? (uint32_t)(j + h) : 0u, o0, o1);
    const bool p0 = act && o0 <= thresh;
    const uint32_t b0 = __ballot_sync(FULL, p0);
    if (p0) {
      const int pos = cnt + __popc(b0 & lt);
      if (pos < cap) cand[pos] = ((uint64_t)o0 << 32) | (uint32_t)j;
    }
    cnt += __popc(b0);
    const bool p1 = has1 && o1 <= thresh;
    const uint32_t b1 = __ballot_sync(FULL, p1);
    if (p1) {
      const int pos = cnt + __popc(b1 & lt);
      if (pos < cap) cand[pos] = ((uint64_t)o1 << 32) | (uint32_t)(j + h);
    }
    cnt += __popc(b1);
  }
  __syncwarp();
  if (force_exact || cnt < nsel || cnt > cap) {
    select_exact(sk0, sk1, n, nsel, out, lane);
    return;
  }
  // rank by counting: composites are unique (the index is), so ranks are too
  for (int ci = lane; ci < cnt; ci += 32) {
    const uint64_t mine = cand[ci];
    int rank = 0;
    for (int j = 0; j < cnt; ++j) rank += (cand[j] < mine) ? 1 : 0;
    if (rank < nsel) out[rank] = (uint16_t)(mine & 0xffffu);
  }
  __syncwarp();
}

}  // namespace rbg
