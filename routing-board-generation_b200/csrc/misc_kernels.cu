// misc_kernels.cu -- key derivation for sharded batches and the on-device
// board validity check.
#include "rbg_host.h"

namespace rbg {

// rows [offset, offset+count) of jax.random.split(key, B):
// flat word f of the split output is o0(f, f+B) for f < B, o1(f-B, f) otherwise.
// Reference convention: keys = split(PRNGKey(0), B) (dataset_generator_jax.py:76),
// per-device contiguous slices (rl_training/setup_train.py:397-399).
__global__ void __launch_bounds__(256) split_keys_kernel(uint32_t k0, uint32_t k1, long long B,
                                                         long long offset, long long count,
                                                         uint32_t *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * count) return;
  const long long f = 2 * offset + i;
  uint32_t o0, o1;
  if (f < B) {
    tf_block(k0, k1, (uint32_t)f, (uint32_t)(f + B), o0, o1);
    out[i] = o0;
  } else {
    tf_block(k0, k1, (uint32_t)(f - B), (uint32_t)f, o0, o1);
    out[i] = o1;
  }
}

int launch_split_keys(uint32_t k0, uint32_t k1, int64_t B, int64_t offset, int64_t count,
                      uint32_t *out, cudaStream_t stream) {
  if (count <= 0) return RBG_OK;
  const int64_t n = 2 * count;
  {
    LaunchScope scope(RBG_K_SPLIT, stream);
    split_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(k0, k1, B, offset, count, out);
  }
  return check_launch("split_keys_kernel");
}

// BoardDatasetGeneratorJAX.__call__ (dataset_generator_jax.py:112-141), one warp per env:
// key, _ = split(key); which = randint(key, (), 0, K) with jax's two-draw formula
// (SURVEY A.9); pins-only grid from heads[which] / targets[which] (heads first, then targets).
constexpr int DS_WARPS = 8;

__global__ void __launch_bounds__(DS_WARPS * 32) dataset_state_kernel(const uint32_t *__restrict__ keys, long long B, int G, int N,
                                                                      const int32_t *__restrict__ heads, const int32_t *__restrict__ targets,
                                                                      uint32_t K, rbg_state st) {
  const int lane = threadIdx.x & 31;
  const long long e = (long long)blockIdx.x * DS_WARPS + (threadIdx.x >> 5);
  if (e >= B) return;
  const int cells = G * G;
  uint32_t k0, k1, b0, b1, h0, h1, l0, l1;
  split2(keys[2 * e], keys[2 * e + 1], k0, k1, b0, b1);  // State.key = split(key)[0]
  split2(k0, k1, h0, h1, l0, l1);                        // randint: k1, k2 = split(key)
  const uint32_t hi = bits_scalar(h0, h1), lo = bits_scalar(l0, l1);
  uint32_t mult = 65536u % K;
  mult = (uint32_t)(((unsigned long long)mult * mult) % K);
  const uint32_t which = ((hi % K) * mult + (lo % K)) % K;  // uint32 wrap-around as in jax
  int32_t *grid = st.grid + e * cells;
  for (int i = lane; i < cells; i += 32) grid[i] = 0;
  int sr = 0, sc = 0, tr = 0, tc = 0;
  if (lane < N) {
    const int32_t *h = heads + (size_t)which * 2 * N, *t = targets + (size_t)which * 2 * N;
    sr = h[lane];
    sc = h[N + lane];
    tr = t[lane];
    tc = t[N + lane];
  }
  __syncwarp();
  if (lane < N && (unsigned)sr < (unsigned)G && (unsigned)sc < (unsigned)G) grid[sr * G + sc] = 3 * lane + POSITION;
  __syncwarp();
  if (lane < N && (unsigned)tr < (unsigned)G && (unsigned)tc < (unsigned)G) grid[tr * G + tc] = 3 * lane + TARGET;
  if (lane < N) {
    st.agent_id[e * N + lane] = lane;
    reinterpret_cast<int2 *>(st.start)[e * N + lane] = make_int2(sr, sc);
    reinterpret_cast<int2 *>(st.target)[e * N + lane] = make_int2(tr, tc);
    reinterpret_cast<int2 *>(st.position)[e * N + lane] = make_int2(sr, sc);
  }
  if (lane == 0) {
    st.step_count[e] = 0;
    st.key[2 * e] = k0;
    st.key[2 * e + 1] = k1;
  }
}

int launch_dataset_state(const uint32_t *keys, int64_t B, int G, int N, const int32_t *heads, const int32_t *targets, int64_t K,
                         const rbg_state &st, cudaStream_t stream) {
  if (B <= 0) return RBG_OK;
  {
    LaunchScope scope(RBG_K_VALIDATE, stream);
    dataset_state_kernel<<<(unsigned)((B + DS_WARPS - 1) / DS_WARPS), DS_WARPS * 32, 0, stream>>>(keys, B, G, N, heads, targets, (uint32_t)K, st);
  }
  return check_launch("dataset_state_kernel");
}

// Board validity, one warp per board (rules: reference
// numpy_implementation/utils/post_processor_utils_numpy.py:34-155 and the
// head->target connectivity of board_processor.py:111-162).
// flags: bit0 encoding range, bit1 head/target count, bit2 neighbour-count
// rule, bit3 head and target not connected, bit4 zero-length wire.
constexpr int VAL_WARPS = 4;

__global__ void __launch_bounds__(VAL_WARPS * 32) validate_kernel(const int32_t *__restrict__ boards, long long B,
                                                                  int G, int N, int32_t *__restrict__ flags) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cells = G * G;
  const int cp = (cells + 15) & ~15;
  uint8_t *grid = smem_raw + (size_t)warp * (2 * cp + 4 * 4 * RBG_MAX_N);
  uint8_t *reached = grid + cp;
  int *heads = reinterpret_cast<int *>(reached + cp);
  int *targets = heads + RBG_MAX_N, *hpos = targets + RBG_MAX_N, *tpos = hpos + RBG_MAX_N;
  for (long long b = (long long)blockIdx.x * VAL_WARPS + warp; b < B; b += (long long)gridDim.x * VAL_WARPS) {
    const int32_t *src = boards + b * cells;
    int fl = 0;
    if (lane < N) heads[lane] = targets[lane] = 0;
    __syncwarp();
    for (int i = lane; i < cells; i += 32) {
      const int v = __ldg(src + i);
      if (v < 0 || v > 3 * N) fl |= 1;
      grid[i] = (uint8_t)v;
      reached[i] = 0;
    }
    fl = __reduce_or_sync(FULL, fl);
    if (fl) {  // the oracle stops at the first rule, so do we
      if (lane == 0) flags[b] = fl;
      __syncwarp();
      continue;
    }
    __syncwarp();
    for (int i = lane; i < cells; i += 32) {
      const int v = grid[i];
      if (v == 0) continue;
      const int w = (v - 1) / 3, t = (v - 1) % 3 + 1;
      if (t == POSITION) {
        atomicAdd(&heads[w], 1);
        hpos[w] = i;
      }
      if (t == TARGET) {
        atomicAdd(&targets[w], 1);
        tpos[w] = i;
      }
    }
    __syncwarp();
    bool eligible = false;
    if (lane < N) {
      if (heads[lane] == 0 && targets[lane] == 1) fl |= 16;
      else if (heads[lane] != 1 || targets[lane] != 1) fl |= 2;
      else eligible = true;
    }
    auto same = [&](int i, int w) { const int v = grid[i]; return v > 0 && (v - 1) / 3 == w; };
    for (int i = lane; i < cells; i += 32) {
      const int v = grid[i];
      if (v == 0) continue;
      const int w = (v - 1) / 3, t = (v - 1) % 3 + 1;
      if (heads[w] == 0) continue;
      const int r = i / G, c = i - r * G;
      int nb = 0;
      nb += (r > 0 && same(i - G, w));
      nb += (r < G - 1 && same(i + G, w));
      nb += (c > 0 && same(i - 1, w));
      nb += (c < G - 1 && same(i + 1, w));
      if (t == PATH ? nb != 2 : nb != 1) fl |= 4;
    }
    // connectivity head -> target through own-wire cells.  When the neighbour-count rule holds
    // (warp-uniform test) every wire cell has at most two same-wire neighbours, so a wire is a simple
    // chain and lane w can just walk it from its head; otherwise flood from every head.
    const bool chains = __reduce_or_sync(FULL, fl & 4) == 0;
    if (chains) {
      if (eligible) {
        int cur = hpos[lane], prev = -1;
        reached[cur] = 1;
        for (int step = 0; step < cells && cur != tpos[lane]; ++step) {
          const int r = cur / G, c = cur - r * G;
          int nxt = -1;
          if (r > 0 && cur - G != prev && same(cur - G, lane)) nxt = cur - G;
          if (r < G - 1 && cur + G != prev && same(cur + G, lane)) nxt = cur + G;
          if (c > 0 && cur - 1 != prev && same(cur - 1, lane)) nxt = cur - 1;
          if (c < G - 1 && cur + 1 != prev && same(cur + 1, lane)) nxt = cur + 1;
          if (nxt < 0) break;
          prev = cur;
          cur = nxt;
          reached[cur] = 1;
        }
      }
      __syncwarp();
    } else {
      if (eligible) reached[hpos[lane]] = 1;
      __syncwarp();
      for (int it = 0; it < cells; ++it) {
        bool changed = false;
        for (int i = lane; i < cells; i += 32) {
          const int v = grid[i];
          if (v == 0 || reached[i]) continue;
          const int w = (v - 1) / 3;
          const int r = i / G, c = i - r * G;
          const bool hit = (r > 0 && reached[i - G] && same(i - G, w)) || (r < G - 1 && reached[i + G] && same(i + G, w)) ||
                           (c > 0 && reached[i - 1] && same(i - 1, w)) || (c < G - 1 && reached[i + 1] && same(i + 1, w));
          if (hit) {
            reached[i] = 1;
            changed = true;
          }
        }
        __syncwarp();
        if (!__any_sync(FULL, changed)) break;
      }
    }
    if (eligible && !reached[tpos[lane]]) fl |= 8;
    fl = __reduce_or_sync(FULL, fl);
    if (lane == 0) flags[b] = fl;
    __syncwarp();
  }
}

int launch_validate(const int32_t *boards, int64_t B, int G, int N, int32_t *flags, cudaStream_t stream) {
  if (B <= 0) return RBG_OK;
  const int cells = G * G, cp = (cells + 15) & ~15;
  const size_t smem = (size_t)VAL_WARPS * (2 * cp + 4 * 4 * RBG_MAX_N);
  int64_t ctas = (B + VAL_WARPS - 1) / VAL_WARPS;
  if (ctas > 148 * 64) ctas = 148 * 64;
  {
    LaunchScope scope(RBG_K_VALIDATE, stream);
    validate_kernel<<<(unsigned)ctas, VAL_WARPS * 32, smem, stream>>>(boards, B, G, N, flags);
  }
  return check_launch("validate_kernel");
}

}  // namespace rbg
