// seqrw_kernel.cu -- SequentialRandomWalkBoard / SequentialRandomWalkGenerator for sm_100a.
//
// Replaces jit(vmap(SequentialRandomWalkBoard.generate)) / generate_starts_ends and
// jit(vmap(SequentialRandomWalkGenerator.__call__)):
//   reference board_generation_methods/jax_implementation/board_generation/sequential_random_walk.py:23-431 (SRW),
//   rl_training/online_generators/sequential_random_walk_generator.py:19-62 (both under
//   /root/reference/routing_board_generation/).
//
// What the reference computes, per board: up to rows + cols attempts, ALL from the same key, with a maximum
// walk length of rows + cols - i in attempt i (SRW:324-392); an attempt places the wires one after the other
// (SRW:290-322): a start cell drawn among the empty cells (pick_start, :34-83), then up to max_length steps,
// each to one of the <= 4 adjacent cells that are free and touch the wire only through the cell it stands on
// (:115-191), until none is left; the attempt succeeds when every wire could start and moved at least once.
// Every draw is jax.random.choice(p in {0,1}, replace=False) = a float32 Gumbel top-k, which for such p is
// "the candidate whose uniform has the largest 23-bit mantissa, lowest index on ties" for any strictly monotone
// float32 log (checked on every draw of the fixtures by tests/tools/make_seqrw_fixtures.py): integer work.
//
// Bound: integer issue (threefry2x32), like prw_kernel.  The work per wire is G*G/2 blocks for the start cell
// and, per step, one split (2 blocks) + the 4 words of the step's draw that can matter (the reference draws
// G + 1 words, one per entry of its padded candidate list; only entries 0..3 can have p = 1).
//
// Organisation: 8 lanes own a board (4 boards per warp), the board in shared memory as uint8 with a 2-cell 0xFF
// border (no bounds tests).  The lanes of a warp run ONE loop whose body is one threefry block per lane plus the
// bookkeeping of whatever phase each board is in, so boards in different phases (different wires, different
// attempts) share every hash pass:
//   SPLIT_WIRE   lanes 0,1: split(key)           -> subkey b             (SRW:307)
//   SPLIT_START  lanes 0,1: split(b)             -> chain key c, draw key d   (:51)
//   PICK         8 blocks of random_bits(d, (G*G,)) per pass, every lane keeps its best empty cell; reduce   (:57-71)
//   SPLIT_WALK   lanes 0,1: split(c)             -> c', step key s        (:204, first step)
//   WALK         lanes 2..5: the step's 4 words under s and the availability of their direction;
//                lanes 0,1: split(c') for the NEXT step in the same pass (the chain does not depend on the
//                board); ballot, pick, move                              (:193-228, :230-288)
// A failed attempt is abandoned at once (its board can only be discarded, :355-366).  Groups take boards from a
// global queue.  The finished board goes to a byte scratch and through se_finish_kernel's outputs (board /
// first POSITION and TARGET cell per wire / State + observation), which SeedExtension shares.
#include "connector_device.cuh"
#include "rbg_host.h"

namespace rbg {

constexpr int SQ_W = 8;
constexpr int SQ_THREADS = 128;
constexpr int SQ_GROUPS = SQ_THREADS / SQ_W;

enum : int { SQ_SPLIT_WIRE = 0, SQ_SPLIT_START = 1, SQ_PICK = 2, SQ_SPLIT_WALK = 3, SQ_WALK = 4, SQ_IDLE = 5, SQ_DONE = 6 };

__global__ void __launch_bounds__(SQ_THREADS) seqrw_walk_kernel(const SeqRwParams p, uint8_t *__restrict__ out_board, uint32_t *__restrict__ out_gkey,
                                                                const int CB, int *__restrict__ queue, const FastDiv divG) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, l = lane & (SQ_W - 1), gbase = lane & ~(SQ_W - 1);
  const unsigned gmask = 0xFFu << gbase;
  const int G = p.G, N = p.N, S = G + 4, cells = G * G;
  const int SB = (S * S + 15) & ~15;
  uint8_t *board = smem_raw + (size_t)(tid / SQ_W) * SB;
  const long long total = p.list ? (long long)(*p.list_count) : p.B;
  const int n1 = cells, h1 = (cells + 1) >> 1;  // pick_start: random_bits(d, (G*G,)), odd sizes padded with a zero counter
  const int n2 = G + 1, h2 = (n2 + 1) >> 1;     // one_step:   random_bits(s, (G+1,))
  const int P = (h1 + SQ_W - 1) / SQ_W;

  int phase = SQ_IDLE;
  long long m = -1, e = -1;
  uint32_t K0 = 0, K1 = 0;        // the board's key (every attempt starts from it)
  uint32_t key0 = 0, key1 = 0;    // the tuple's key between wires
  uint32_t c0 = 0, c1 = 0;        // chain key of the walk (committed)
  uint32_t cn0 = 0, cn1 = 0;      // split(c)[0], ready for when the step is taken
  uint32_t s0 = 0, s1 = 0;        // SPLIT_START: subkey b; WALK: the step's draw key
  uint32_t d0 = 0, d1 = 0;        // pick_start's draw key
  int w = 0, L = 0, t = 0, filled = 0, cur = 0, startc = 0, j = 0, steps = 0, attempt = 0;
  uint32_t bestm = 0;
  int besti = -1;

  for (;;) {
    __syncwarp();
    if (__all_sync(FULL, phase == SQ_DONE)) break;
    bool fresh = false;  // (re)start an attempt: empty board, wire 0
    if (phase == SQ_IDLE) {
      int nm = 0;
      if (l == 0) nm = atomicAdd(queue, 1);
      nm = __shfl_sync(gmask, nm, gbase);
      if ((long long)nm >= total) {
        phase = SQ_DONE;
      } else {
        m = nm;
        e = p.list ? (long long)p.list[m] : m;
        uint32_t k0 = p.keys[2 * e], k1 = p.keys[2 * e + 1], a0, a1, b0, b1;
        for (int sp = 0; sp < p.extra_split; ++sp) {  // generator / auto-reset: key = split(key)[0]
          split2(k0, k1, a0, a1, b0, b1);
          k0 = a0;
          k1 = a1;
        }
        K0 = k0;
        K1 = k1;
        if (l == 0) {
          out_gkey[2 * m] = k0;
          out_gkey[2 * m + 1] = k1;
        }
        L = 2 * G - 1;  // SRW:352 max_length_int - i, i = 1
        attempt = 1;
        fresh = true;
      }
    }
    if (fresh) {
      for (int r = 0; r < S; ++r) {
        const bool rin = r >= 2 && r < G + 2;
        for (int c = l; c < S; c += SQ_W) board[r * S + c] = (rin && c >= 2 && c < G + 2) ? 0 : 0xFF;
      }
      w = 0;
      filled = 0;
      steps = 0;
      key0 = K0;
      key1 = K1;
      phase = SQ_SPLIT_WIRE;
      __syncwarp(gmask);
    }

    // ---- one threefry block per lane
    uint32_t hk0 = key0, hk1 = key1, x0 = (uint32_t)(l & 1), x1 = (uint32_t)(l & 1) + 2u;  // split(): blocks (0,2), (1,3)
    if (phase == SQ_SPLIT_START) {
      hk0 = s0;
      hk1 = s1;
    } else if (phase == SQ_PICK) {
      const int i = j * SQ_W + l;
      hk0 = d0;
      hk1 = d1;
      x0 = (uint32_t)i;
      x1 = (i + h1 < n1) ? (uint32_t)(i + h1) : 0u;
    } else if (phase == SQ_SPLIT_WALK) {
      hk0 = c0;
      hk1 = c1;
    } else if (phase == SQ_WALK) {
      if (l >= 2 && l < 6) {
        const int k = l - 2, blk = k < h2 ? k : k - h2;  // word k of random_bits(s, (G+1,)): o0 of block k, or o1 of block k - h2
        hk0 = s0;
        hk1 = s1;
        x0 = (uint32_t)blk;
        x1 = (blk + h2 < n2) ? (uint32_t)(blk + h2) : 0u;
      } else {
        hk0 = cn0;
        hk1 = cn1;
      }
    }
    __syncwarp();
    uint32_t o0, o1;
    tf_block(hk0, hk1, x0, x1, o0, o1);
    const uint32_t a0 = __shfl_sync(FULL, o0, gbase), a1 = __shfl_sync(FULL, o0, gbase + 1);
    const uint32_t b0 = __shfl_sync(FULL, o1, gbase), b1 = __shfl_sync(FULL, o1, gbase + 1);

    // ---- what the pass meant for this board
    bool failed = false, finished = false;
    if (phase == SQ_SPLIT_WIRE) {
      s0 = b0;  // SRW:307 subkey; the other half is dropped (:306 takes the key from the tuple)
      s1 = b1;
      phase = SQ_SPLIT_START;
    } else if (phase == SQ_SPLIT_START) {
      c0 = a0;  // SRW:51 key (the tuple's key from here on), subkey
      c1 = a1;
      d0 = b0;
      d1 = b1;
      if (filled == cells) {  // SRW:53 can_start False: the attempt cannot succeed
        failed = true;
      } else {
        phase = SQ_PICK;
        j = 0;
        bestm = 0;
        besti = -1;
      }
    } else if (phase == SQ_PICK) {
      const int i = j * SQ_W + l;
      if (i < h1) {
        uint32_t q, r;
        divG.divmod((uint32_t)i, q, r);
        if (board[(q + 2) * S + r + 2] == 0) {
          const uint32_t mm = o0 >> 9;
          if (besti < 0 || mm > bestm || (mm == bestm && i < besti)) bestm = mm, besti = i;
        }
        const int i2 = i + h1;
        if (i2 < n1) {
          divG.divmod((uint32_t)i2, q, r);
          if (board[(q + 2) * S + r + 2] == 0) {
            const uint32_t mm = o1 >> 9;
            if (besti < 0 || mm > bestm || (mm == bestm && i2 < besti)) bestm = mm, besti = i2;
          }
        }
      }
      if (++j == P) {
        for (int off = SQ_W / 2; off > 0; off >>= 1) {
          const uint32_t om = __shfl_xor_sync(gmask, bestm, off);
          const int oi = __shfl_xor_sync(gmask, besti, off);
          if (oi >= 0 && (besti < 0 || om > bestm || (om == bestm && oi < besti))) bestm = om, besti = oi;
        }
        uint32_t q, r;
        divG.divmod((uint32_t)besti, q, r);  // SRW:66 divmod(flat, rows)
        startc = (int)((q + 2) * S + r + 2);
        cur = startc;
        if (l == 0) board[startc] = (uint8_t)(3 * w + POSITION);  // SRW:71
        filled++;
        t = 0;
        phase = SQ_SPLIT_WALK;
        __syncwarp(gmask);
      }
    } else if (phase == SQ_SPLIT_WALK) {
      cn0 = a0;  // SRW:204 of the first step
      cn1 = a1;
      s0 = b0;
      s1 = b1;
      phase = SQ_WALK;
    } else if (phase == SQ_WALK) {
      // SRW:115-140 available_cells of the cell the wire stands on: [up, down, left, right]
      bool av = false;
      uint32_t mant = 0;
      if (l >= 2 && l < 6 && t < L) {
        const int k = l - 2;
        const int cand = cur + (k == 0 ? -S : k == 1 ? S : k == 2 ? -1 : 1);
        if (board[cand] == 0) {  // :142-156 is_cell_free (the border is never free)
          const uint32_t base = 3u * w + 1u;  // :158-191 touches the wire through at most one cell
          const int touching = (int)own_wire(board[cand - S], base) + (int)own_wire(board[cand + S], base) + (int)own_wire(board[cand - 1], base) +
                               (int)own_wire(board[cand + 1], base);
          av = touching <= 1;
        }
        mant = (k < h2 ? o0 : o1) >> 9;
      }
      const unsigned bal = (__ballot_sync(gmask, av) >> (gbase + 2)) & 0xFu;
      if (bal == 0) {  // :241 can_step False (or max_length steps done): the walk is over
        if (t == 0) {
          failed = true;  // :319 the wire did not move
        } else {
          if (l == 0) board[startc] = (uint8_t)(3 * w + POSITION);  // :286
          key0 = c0;
          key1 = c1;
          ++w;
          if (w == N)
            finished = true;
          else
            phase = SQ_SPLIT_WIRE;
          __syncwarp(gmask);
        }
      } else {
        // :211-217 choice over the padded list: largest mantissa among the available entries, lowest index on ties
        int pick = -1;
        uint32_t pm = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t mk = __shfl_sync(gmask, mant, gbase + 2 + k);
          if (((bal >> k) & 1u) && (pick < 0 || mk > pm)) pick = k, pm = mk;
        }
        const int nxt = cur + (pick == 0 ? -S : pick == 1 ? S : pick == 2 ? -1 : 1);
        if (l == 0) {
          board[nxt] = (uint8_t)(3 * w + TARGET);  // :221
          board[cur] = (uint8_t)(3 * w + PATH);    // :223-227
        }
        cur = nxt;
        ++t;
        ++filled;
        ++steps;
        c0 = cn0;  // the step is taken: its split is the one computed a pass ago; this pass computed the next one
        c1 = cn1;
        cn0 = a0;
        cn1 = a1;
        s0 = b0;
        s1 = b1;
        __syncwarp(gmask);
      }
    }

    if (failed) {  // SRW:342-366: the next attempt walks one step less, from the same key
      --L;
      ++attempt;
      if (L <= 0) {  // an attempt with max_length 0 cannot move: SRW:389-391 zero board
        attempt = 0;
        steps = 0;
        finished = true;
      } else {
        for (int r = 2; r < G + 2; ++r)
          for (int c = 2 + l; c < G + 2; c += SQ_W) board[r * S + c] = 0;
        w = 0;
        filled = 0;
        steps = 0;
        key0 = K0;
        key1 = K1;
        phase = SQ_SPLIT_WIRE;
        __syncwarp(gmask);
      }
    }
    if (finished) {
      uint32_t *ob = reinterpret_cast<uint32_t *>(out_board + (size_t)m * CB);
      for (int qd = l; qd < (CB >> 2); qd += SQ_W) {
        uint32_t word = 0;
        if (attempt != 0) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int i = 4 * qd + b;
            if (i < cells) {
              uint32_t q, r;
              divG.divmod((uint32_t)i, q, r);
              word |= (uint32_t)board[(q + 2) * S + r + 2] << (8 * b);
            }
          }
        }
        ob[qd] = word;
      }
      if (l == 0 && p.stats) {
        p.stats[2 * e] = attempt;
        p.stats[2 * e + 1] = steps;
      }
      phase = SQ_IDLE;
      __syncwarp(gmask);
    }
  }
}

int launch_seqrw(SeqRwParams p, int64_t max_boards, cudaStream_t stream) {
  const int G = p.G, N = p.N;
  if (G < 3) return set_error(RBG_EINVAL, "SequentialRandomWalk: rows=%d (available_cells pads with jnp.full(rows - 3, -1): rows >= 3)", G);
  const int S = G + 4, SB = (S * S + 15) & ~15, CB = (G * G + 15) & ~15;
  const size_t smem = (size_t)SQ_GROUPS * SB;
  const size_t n = (size_t)max_boards;
  const size_t o_gkey = (n * CB + 255) & ~(size_t)255, o_queue = o_gkey + ((n * 8 + 255) & ~(size_t)255), total = o_queue + 256;
  uint8_t *base = nullptr;
  cudaError_t ce = cudaMallocAsync(reinterpret_cast<void **>(&base), total, stream);
  if (ce != cudaSuccess) return set_cuda_error(ce, "cudaMallocAsync(SequentialRandomWalk scratch)");
  uint32_t *gkey = reinterpret_cast<uint32_t *>(base + o_gkey);
  int *queue = reinterpret_cast<int *>(base + o_queue);
  int rc = RBG_OK;
  do {
    if ((ce = cudaMemsetAsync(queue, 0, 256, stream)) != cudaSuccess) {
      rc = set_cuda_error(ce, "cudaMemsetAsync(SequentialRandomWalk queue)");
      break;
    }
    if (smem > 48 * 1024 && (ce = cudaFuncSetAttribute(seqrw_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) {
      rc = set_cuda_error(ce, "seqrw_walk_kernel shared memory");
      break;
    }
    int per_sm = 0;
    if ((ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seqrw_walk_kernel, SQ_THREADS, smem)) != cudaSuccess) {
      rc = set_cuda_error(ce, "cudaOccupancyMaxActiveBlocksPerMultiprocessor(seqrw_walk_kernel)");
      break;
    }
    const long long resident = (long long)(per_sm < 1 ? 1 : per_sm) * device_sm_count();
    const long long need = (max_boards + SQ_GROUPS - 1) / SQ_GROUPS;
    const unsigned ctas = (unsigned)(need < resident ? (need < 1 ? 1 : need) : resident);
    {
      LaunchScope scope(RBG_K_SEQRW, stream);
      seqrw_walk_kernel<<<ctas, SQ_THREADS, smem, stream>>>(p, base, gkey, CB, queue, FastDiv::make((uint32_t)G));
    }
    if ((rc = check_launch("seqrw_walk_kernel"))) break;
    SeedExtParams f;
    memset(&f, 0, sizeof(f));
    f.keys = p.keys;
    f.B = p.B;
    f.G = G;
    f.N = N;
    f.mode = p.mode;
    f.solved = p.board;
    f.solved_f32 = p.float_board;
    f.starts = p.starts;
    f.ends = p.ends;
    f.st = p.st;
    f.ts = p.ts;
    f.observe = p.observe;
    f.list = p.list;
    f.list_count = p.list_count;
    rc = launch_board_finish(f, base, gkey, CB, max_boards, RBG_K_SEQRW, stream);
  } while (0);
  cudaFreeAsync(base, stream);
  return rc;
}

}  // namespace rbg
