/*
 * rbg_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the reference's algorithms for the hot path
 * (ParallelRandomWalk, SeedExtension, Uniform generator, Jumanji Connector
 * reset/step) plus the jax.random (jax==0.4.8, threefry2x32,
 * jax_threefry_partitionable=False) primitives they call.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product path (routing-board-generation_b200/) never
 * links, imports or falls back to it.
 *
 * Parity status (see DESIGN.md "Oracle pinning"):
 *   - jax.random + ParallelRandomWalk + Uniform generator: PINNED by the
 *     reference's own golden vectors (test_parallel_random_walk_board.py and
 *     the stored notebook outputs), see tests/golden/reference_goldens.json.
 *   - ParallelRandomWalk collision branch, SeedExtension, Connector
 *     step/reset/observation/reward/extras: PARITY UNPINNED by reference
 *     goldens (none exist); checked instead by running the reference's own
 *     Python files under a NumPy stand-in for jax (tests/tools/jax_shim) where
 *     the source is in /root/reference, and restated from jumanji==0.2.2's
 *     published source where it is not (Connector).
 *
 * All citations are path:line relative to /root/reference/.
 * PRW  = routing_board_generation/board_generation_methods/jax_implementation/board_generation/parallel_random_walk.py
 * SE   = .../jax_implementation/board_generation/seed_extension.py
 * PPU  = .../jax_implementation/utils/post_processor_utils_jax.py
 * GU   = .../jax_implementation/utils/grid_utils.py
 * PRWG = routing_board_generation/rl_training/online_generators/parallel_random_walk_generator.py
 * UG   = routing_board_generation/rl_training/online_generators/uniform_generator.py
 * RSG  = routing_board_generation/rl_training/online_generators/random_seed_generator.py
 */
#ifndef RBG_ORACLE_H
#define RBG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_G 64
#define ORC_MAX_N 85

/* generator kinds for orc_state / orc_connector_reset */
#define ORC_GEN_PRW 0
#define ORC_GEN_UNIFORM 1
#define ORC_GEN_SEEDEXT 2
#define ORC_GEN_SEQRW 4 /* 3 = dataset (orc_dataset_state) */

/* ---- jax.random (0.4.8) primitives ------------------------------------- */
void orc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1,
                      uint32_t *o0, uint32_t *o1);
void orc_prng_key(uint64_t seed, uint32_t key[2]);
void orc_split(const uint32_t key[2], int num, uint32_t *out /* [num,2] */);
void orc_random_bits(const uint32_t key[2], int n, uint32_t *out /* [n] */);
float orc_uniform(const uint32_t key[2]);
void orc_shuffle_iota(const uint32_t key[2], int n, int32_t *out /* [n] */);
int32_t orc_randint(const uint32_t key[2], int32_t lo, int32_t hi);
void orc_randint_vec(const uint32_t key[2], int n, int32_t lo, int32_t hi,
                     int32_t *out);
int32_t orc_choice_p4(const uint32_t key[2], const int32_t a[4],
                      const uint8_t p[4]);

/* ---- ParallelRandomWalk helpers exposed for the reference's unit goldens */
void orc_prw_adjacent_cells(int G, int cell, int32_t out[4]);
void orc_prw_available_cells(int G, const int32_t *grid, int cell,
                             int32_t out[4]);
int orc_prw_is_cell_free(int G, const int32_t *grid, int cell);
int orc_prw_action_from_positions(int G, int p1, int p2);
void orc_prw_initialise_agents(const uint32_t key[2], int G, int N,
                               int32_t *grid, int32_t *pos /* [N,2] */);
/* one _step: returns number of collided agents; actions_out may be NULL */
int orc_prw_step(const uint32_t key[2], int G, int N, int32_t *grid,
                 int32_t *pos /* [N,2] */, int32_t *actions_out,
                 uint32_t next_key[2]);
int orc_prw_continue_stepping(int G, int N, const int32_t *grid,
                              const int32_t *pos);

/* generate_board: heads[2,N], targets[2,N], solved[G,G]; stats may be NULL
 * stats[0]=while-loop trips, stats[1]=collided agent-moves. returns 0 / <0 */
int orc_prw_generate(const uint32_t key[2], int G, int N, int32_t *heads,
                     int32_t *targets, int32_t *solved, int32_t *stats);

/* ---- SeedExtension ------------------------------------------------------- */
int orc_seedext_seeded_board(const uint32_t key[2], int G, int N,
                             int32_t *board);
/* stats (may be NULL): [0]=extension sweeps, [1]=bfs pops total, [2]=degenerate flag */
int orc_seedext_solved(const uint32_t key[2], int G, int N, float randomness,
                       int two_sided, int iterations, int64_t ext_steps,
                       int32_t *board, int32_t *stats);
int orc_seedext_starts_ends(const uint32_t key[2], int G, int N,
                            float randomness, int two_sided, int iterations,
                            int64_t ext_steps, int32_t *starts /* [2,N] */,
                            int32_t *ends /* [2,N] */);
void orc_extend_wires(int G, int32_t *board, const uint32_t key[2],
                      float randomness, int two_sided, int64_t ext_steps,
                      int32_t *sweeps_out);
int orc_optimise_wire(const uint32_t key[2], int G, int32_t *board, int wire,
                      int32_t *pops_out);

/* ---- SequentialRandomWalkBoard (SRW = .../jax_implementation/board_generation/sequential_random_walk.py;
 *      pinned by tests/golden/seqrw_reference.json: the reference's own class run on the NumPy stand-in) ---- */
/* generate SRW:324-392: board[G,G] (the reference returns the same codes as float32); stats may be NULL:
 * [0] = attempt that succeeded (1 .. 2G, 0 = none: zero board), [1] = steps of that attempt */
int orc_seqrw_generate(const uint32_t key[2], int G, int N, int32_t *board, int32_t *stats);
/* generate_starts_ends SRW:394-431 -> starts[2,N], ends[2,N] */
int orc_seqrw_starts_ends(const uint32_t key[2], int G, int N, int32_t *starts, int32_t *ends);
int orc_seqrw_generate_batch(const uint32_t *keys, int64_t B, int G, int N, int32_t *boards,
                             int32_t *stats /* [B,2] or NULL */, int nthreads);

/* ---- Generator __call__(key) -> State ---------------------------------- */
/* State fields (one env): grid[G,G], step_count, agent_id[N], start[N,2],
 * target[N,2], position[N,2], key_out[2] */
int orc_state(int kind, const uint32_t key[2], int G, int N, int32_t *grid,
              int32_t *step_count, int32_t *agent_id, int32_t *start,
              int32_t *target, int32_t *position, uint32_t key_out[2]);

/* ---- Connector (jumanji==0.2.2) ---------------------------------------- */
void orc_connector_action_mask(int G, int N, const int32_t *grid,
                               const int32_t *target, const int32_t *position,
                               uint8_t *mask /* [N,5] */);
void orc_connector_obs(int G, int N, const int32_t *grid,
                       int32_t *obs /* [N,G,G] */);
void orc_connector_extras(int G, int N, const int32_t *grid,
                          const int32_t *target, const int32_t *position,
                          int32_t *num_connections, float *ratio_connections,
                          int32_t *total_path_length);
/* step: grid/step_count/position updated in place */
void orc_connector_step(int G, int N, int32_t *grid, int32_t *step_count,
                        const int32_t *target, int32_t *position,
                        const int32_t *action, int time_limit,
                        float timestep_reward, float connected_reward,
                        int32_t *obs, uint8_t *mask, float *reward,
                        float *discount, int8_t *step_type,
                        int32_t *num_connections, float *ratio_connections,
                        int32_t *total_path_length);

/* ---- validity: numpy_implementation/utils/post_processor_utils_numpy.py:34-155 (UP),
 *      numpy_implementation/utils/board_processor.py:111-162,391-487 (BP); pinned by
 *      tests/golden/validity_reference.npz (the reference's own verdicts) ---- */
/* returns a bitmask, 0 = valid.  1 EncodingOutOfRangeError (nothing else evaluated), 2 MissingHeadTailError,
 * 4 InvalidWireStructureError (verify_wire_validity False), 8 head / target not connected through own cells,
 * 16 zero-length wire (lone TARGET), 32 duplicated head / target, 64 PathNotFoundError (no path through
 * own + EMPTY cells), 128 rule 2 or 4 broken by something other than a zero-length wire */
int orc_validate_board(int G, int N, const int32_t *board);

/* ---- batched (OpenMP over boards/envs) ---------------------------------- */
void orc_split_batch_slice(const uint32_t key[2], int64_t B, int64_t offset,
                           int64_t count, uint32_t *out /* [count,2] */);
int orc_prw_generate_batch(const uint32_t *keys, int64_t B, int G, int N,
                           int32_t *heads, int32_t *targets, int32_t *solved,
                           int32_t *stats /* [B,2] or NULL */, int nthreads);
int orc_seedext_solved_batch(const uint32_t *keys, int64_t B, int G, int N,
                             float randomness, int two_sided, int iterations,
                             int64_t ext_steps, int32_t *boards,
                             int32_t *stats /* [B,3] or NULL */, int nthreads);
int orc_state_batch(int kind, const uint32_t *keys, int64_t B, int G, int N,
                    int32_t *grid, int32_t *step_count, int32_t *agent_id,
                    int32_t *start, int32_t *target, int32_t *position,
                    uint32_t *key_out, int nthreads);
void orc_connector_observe_batch(int64_t B, int G, int N, const int32_t *grid,
                                 const int32_t *target, const int32_t *position,
                                 int32_t *obs, uint8_t *mask,
                                 int32_t *num_connections,
                                 float *ratio_connections,
                                 int32_t *total_path_length, int nthreads);
/* autoreset_kind < 0: plain Connector.step; else VmapAutoResetWrapper with
 * that generator kind (state.key consumed/updated) */
void orc_connector_step_batch(int64_t B, int G, int N, int32_t *grid,
                              int32_t *step_count, int32_t *start,
                              int32_t *target, int32_t *position,
                              uint32_t *key, const int32_t *action,
                              int time_limit, float timestep_reward,
                              float connected_reward, int autoreset_kind,
                              int32_t *obs, uint8_t *mask, float *reward,
                              float *discount, int8_t *step_type,
                              int32_t *num_connections,
                              float *ratio_connections,
                              int32_t *total_path_length,
                              int32_t *obs_step_count, int nthreads);
void orc_connector_step_batch_ds(int64_t B, int G, int N, int32_t *grid,
                                 int32_t *step_count, int32_t *start,
                                 int32_t *target, int32_t *position,
                                 uint32_t *key, const int32_t *action,
                                 int time_limit, float timestep_reward,
                                 float connected_reward, int autoreset_kind /* 3 = dataset */,
                                 const int32_t *ds_heads, const int32_t *ds_targets, int64_t ds_K,
                                 int32_t *obs, uint8_t *mask, float *reward,
                                 float *discount, int8_t *step_type,
                                 int32_t *num_connections,
                                 float *ratio_connections,
                                 int32_t *total_path_length,
                                 int32_t *obs_step_count, int nthreads);
/* EvaluateEmptyBoard (benchmarking/benchmarks/empty_board_evaluation.py:31-155): scored board, count_detours,
 * heatmap_score_diversity */
void orc_board_statistics(int G, const int32_t *board, int count_current_wire, int32_t *scored, int32_t *detours,
                          int32_t *diversity);
void orc_board_statistics_batch(int64_t B, int G, const int32_t *boards, int count_current_wire, int32_t *scored,
                                int32_t *detours, int32_t *diversity, int nthreads);
void orc_validate_batch(int64_t B, int G, int N, const int32_t *boards,
                        int32_t *flags, int nthreads);
/* the bench's random policy (OUR convention, not a parity surface):
 * see rbg_random_actions in include/rbg_b200.h */
void orc_random_actions_batch(int64_t B, int G, int N, const int32_t *grid,
                              const int32_t *target, const int32_t *position,
                              const int32_t *step_count, const uint32_t *key,
                              int32_t *action, int nthreads);
int orc_max_threads(void);

/* BoardDatasetGeneratorJAX.__call__: heads / targets are [K,2,N] */
int orc_dataset_state(const uint32_t key_in[2], int G, int N, const int32_t *heads,
                      const int32_t *targets, int64_t K, int32_t *grid,
                      int32_t *step_count, int32_t *agent_id, int32_t *start,
                      int32_t *target, int32_t *position, uint32_t key_out[2]);

#ifdef __cplusplus
}
#endif
#endif
