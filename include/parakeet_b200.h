/*
 * parakeet_b200.h -- C ABI of libparakeet_b200.so, the B200 (sm_100a) FastSLAM 1.0 hot path.
 *
 * This is the drop-in boundary for the particle-filter path of buckbaskin/parakeet_slam
 * (reference file src/prkt_core_v2.py).  The reference is pure Python with no FFI of its
 * own, so each entry point names the reference method whose per-particle loop it replaces
 * (file:line relative to /root/reference/src).  The Python binding a maintainer would add is
 * a ctypes stub -- see INTEGRATION.md and parakeet_slam_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends
 *     in _host; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - the library never allocates or frees persistent device memory behind the caller's back: the
 *     caller owns all state (PyTorch tensors on the Python side) and passes a workspace where one
 *     is needed.  The one exception is explicit: pk_peer_alloc / pk_peer_free, because memory that
 *     is shared with other ranks through CUDA IPC must be a plain cudaMalloc allocation
 *   - every function returns 0 on success or a negative PK_E* code; pk_last_error() gives a
 *     thread-local message.  No C++ exception crosses the boundary.
 *   - all work is asynchronous on `stream`; nothing here synchronises the device
 *
 * Device data layout (see DESIGN.md "Data layout in HBM")
 *   pose4   double[M][4]    x, y, heading (as read back through the quaternion), weight
 *   aux2    int[M][2]       n_live landmarks, next_id            (travels with the particle)
 *   slot    int[M]          index of the particle's landmark block in the pool
 *   pool    bytes           n_slots blocks of pk_block_bytes(capacity, dtype); one block =
 *                           [hot region: capacity x 4 B, padded to 64 B][cold region: capacity x COLD]
 *                             hot  4 B  = colour KEY: r,g,b rounded and clamped to bytes (the only
 *                                         part streamed for every landmark by the fused kernel)
 *                             cold f32 64 B  = r,g,b, x,y, Sp lower triangle (3), Sc lower triangle (6)
 *                                              as float, id, meta  (one 64-byte DRAM granule)
 *                                  f64 160 B = r,g,b, x,y, Sp[2][2], Sc[3][3] as double, id, meta, pad
 *                           meta = update_count | PK_META_IMMUTABLE | PK_META_POTENTIAL
 *                           The 5x5 landmark covariance of the reference is stored as its two
 *                           diagonal blocks; the cross blocks are exactly zero in every state the
 *                           reference can reach (H has exact zeros, prkt_core_v2.py:799-802).
 */
#ifndef PARAKEET_B200_H
#define PARAKEET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_ABI_VERSION 1

#define PK_OK 0
#define PK_EINVAL (-1)   /* bad argument (null pointer, size, dtype, K > PK_MAX_OBS ...) */
#define PK_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define PK_EARCH (-3)    /* device is not sm_100 */

#define PK_DTYPE_F32 0   /* landmark storage fp32 (arithmetic is fp64 either way) */
#define PK_DTYPE_F64 1   /* landmark storage fp64: the parity instantiation */

/* Spawn mode: every `dtype` argument is a layout code -- the storage type in the low
 * byte and the number of orphan-reading slots per particle above it.  The orphan region
 * [64-byte header: int total | n x 64-byte readings: double x, y, cos(ray), sin(ray), r, g, b, id]
 * follows the cold region inside the particle's block and travels with it. */
#define PK_MAX_ORPHANS 1024
/* Arithmetic flag of the layout code (pk_measurement_update only; needs PK_DTYPE_F32 storage): the landmark
 * algebra of K2 -- gates, Mahalanobis forms, EKF gain and covariance update -- runs in fp32 on the fp32 records;
 * poses, pose-landmark differences, the importance weight (its exp and the scan-order product) and everything
 * in resampling stay fp64.  The match / no-match decision (fp64 underflow of the likelihood) is taken
 * from the pdf exponents.  Throughput mode: >= 90 % of the reference's indices, state to fp32 rounding. */
#define PK_DTYPE_ARITH_F32 0x1000000
#define PK_DTYPE_WITH_ORPHANS(base, n) ((base) | ((n) << 8))

#define PK_MAX_OBS 64        /* blobs per frame handled by one pk_measurement_update */
#define PK_SCAN_BLOCK 1024   /* particles per weight-scan block (fixed: results must not depend on the shard count) */

#define PK_META_COUNT_MASK 0x00ffffff
#define PK_META_IMMUTABLE  0x10000000   /* Feature.__immutable__   (prkt_core_v2.py:883, prkt_ros.py:41) */
#define PK_META_POTENTIAL  0x20000000   /* lives in potential_features, id < 0 (prkt_core_v2.py:287, 679) */

/* indices into the stats array written by pk_measurement_update (uint64 each) */
#define PK_STAT_MATCHED 0        /* (particle, blob) pairs with id != 0 */
#define PK_STAT_UNMATCHED 1      /* pairs with id == 0 (orphaned readings, prkt_core_v2.py:92-95) */
#define PK_STAT_EVALUATED 2      /* exact likelihood evaluations (survivors of the colour pre-filter) */
#define PK_STAT_FLAGS 3          /* OR of PK_FLAG_* */
#define PK_STAT_SAME_LANDMARK 4  /* updates that had to wait for an earlier blob on the same landmark */
#define PK_STAT_PROMOTED 5       /* potential -> full promotions (prkt_core_v2.py:114-118) */
#define PK_STAT_SPAWNED 6        /* potential landmarks created by pk_spawn_update (add_new_feature :653-680) */
#define PK_STAT_ORPHANED 7       /* readings stored by pk_spawn_update (add_orphaned_reading :740-746) */
#define PK_NUM_STATS 8

/* peer (NVLink) exchange of the sharded filter */
#define PK_MAX_RANKS 32          /* ranks of one node sharing a filter (lane g of a warp serves rank g) */
#define PK_XPLAN_LONGS 80        /* int64 words of the device-resident exchange plan */
#define PK_PEER_HANDLE_BYTES 64  /* size of an exported peer-memory handle (cudaIpcMemHandle_t) */
#define PK_PEER_OVERFLOW 1ull    /* status bit: a rank would receive more than the exchange capacity */
#define PK_PEER_TIMEOUT 2ull     /* status bit: a peer did not reach a barrier within the time-out */
#define PK_PEER_STATUS_WORDS 4   /* status array of the fused entry points (pk_resample_thresholds_peer,
                                  * pk_resample_gather_peer): [0] sticky PK_PEER_* bits, [1] / [2] nanoseconds this rank
                                  * has spent inside the first / second flag barrier of its frames, [3] frames counted */
/* words of the exchange plan a caller may read back (the rest is internal) */
#define PK_XP_EMIT_LO 0
#define PK_XP_EMIT_N 1
#define PK_XP_N_LO 2
#define PK_XP_N_LOC 3
#define PK_XP_N_HI 4
#define PK_XP_N_BELOW 5
#define PK_XP_N_ABOVE 6
#define PK_XP_ABOVE_START 7
#define PK_XP_N_SEND 8
#define PK_XP_N_IN 9
#define PK_XP_OVERFLOW 10
#define PK_XP_RANK_LO 16
#define PK_XP_RANK_LOC (16 + PK_MAX_RANKS)

#define PK_FLAG_SINGULAR_COV 1u      /* a landmark covariance block had det <= 0 (SciPy would raise) */
#define PK_FLAG_NONFINITE_WEIGHT 2u  /* a particle weight became NaN/Inf */
#define PK_FLAG_REPROMOTED 4u        /* a blob hit a landmark promoted earlier in the same frame
                                        (the reference raises KeyError there, prkt_core_v2.py:98,312) */
#define PK_FLAG_MAP_FULL 8u          /* spawn mode: a new landmark did not fit the particle's capacity (dropped) */
#define PK_FLAG_ORPHAN_EXPIRED 16u   /* spawn mode: the orphan ring overwrote its oldest reading (the reference
                                        keeps every reading for ever; a reported deviation) */
#define PK_FLAG_SPAWN_DEGENERATE 32u /* spawn mode: cross_readings found parallel rays (the reference raises
                                        TypeError there, prkt_core_v2.py:665-667, 735-736); reading orphaned */

/* Literals of the reference, gathered in one struct (SURVEY.md section 5 "Config / flags"). */
typedef struct pk_params {
    double bearing_gate;     /* 0.5 rad      prkt_core_v2.py:433 */
    double position_gate;    /* pi/2         prkt_core_v2.py:474 */
    double color_gate;       /* 300          prkt_core_v2.py:441 */
    double no_match_weight;  /* 0.1          prkt_core_v2.py:857 */
    double qt_diag;          /* 0.1          prkt_core_v2.py:50-53 (Qt = qt_diag * I4) */
    int promote_count;       /* 5            prkt_core_v2.py:114 (promote when update_count > 5) */
    int model;               /* 0 = the reference's measurement model, as written (default); PK_MODEL_TEXTBOOK */
} pk_params;
/* PK_MODEL_TEXTBOOK: the EKF update with the textbook bearing model instead of the reference's (SURVEY.md finding F4,
 * section 8(f) row 3): predicted bearing in the ROBOT frame, atan2(dy, dx) - heading (the reference predicts the
 * world-frame bearing, prkt_core_v2.py:871); Jacobian row [-dy/q, +dx/q] (the reference writes [+dy/q, +dx/q],
 * :785-797); bearing innovation wrapped to (-pi, pi] (the reference does not wrap, :846, :911).  Association
 * (probability_of_match) is unchanged.  Not reference behaviour: no reference fixture exists for it; pinned against
 * the NumPy restatement run with the same switch. */
#define PK_MODEL_TEXTBOOK 1
/* PK_MODEL_LOG_WEIGHTS: the importance factors and the particle weight are carried as LOGARITHMS (pose4[.][3] = sum of
 * the log factors; an unseen blob adds log(no_match_weight)), so a frame of very unlikely observations cannot
 * underflow every weight to zero as the reference's fp64 product can (finding F3 applies to the association only).
 * pk_log_weights_max / pk_log_weights_normalise turn them into linear weights exp(lw - max) (largest weight 1)
 * before the resampling scan; across shards the max and the sums are all-reduced (NCCL) in between. */
#define PK_MODEL_LOG_WEIGHTS 2

int pk_version(void);
const char* pk_last_error(void);
int pk_default_params(pk_params* out);
/* 0 if the current device is sm_100 (B200); PK_EARCH otherwise, PK_ECUDA if there is none. */
int pk_check_device(void);

/* ---- layout ------------------------------------------------------------------------------ */
int pk_hot_bytes(int dtype);
int pk_cold_bytes(int dtype);
long long pk_block_bytes(int capacity, int dtype);

/* ---- construction: FastSLAM.__init__ / FilterParticle.__init__ / load_feature_list
 *      (prkt_core_v2.py:38-57, 279-299) -------------------------------------------------------- */
/* pose <- (0,0,0), weight <- 1, slot[i] <- i, aux <- (n_live, next_id) */
int pk_init_particles(double* pose4, int* slot, int* aux2, long long M, int n_live, int next_id,
                      void* stream);
/* Write one map (n landmarks given as fp64 SoA device arrays mean5[n][5], covp[n][4],
 * covc[n][9], meta[n], ids[n]) into blocks [slot_lo, slot_hi) of the pool. */
int pk_map_broadcast(void* pool, int capacity, int dtype, long long slot_lo, long long slot_hi,
                     int n, const double* mean5, const double* covp, const double* covc,
                     const int* meta, const int* ids, void* stream);
/* Per-particle maps <-> fp64 SoA arrays shaped [count][capacity][...] for particles
 * [p_lo, p_lo+count) (particles[i].feature_set views, tests, checkpoints). */
int pk_map_export(const void* pool, int capacity, int dtype, const int* slot, long long p_lo,
                  long long count, double* mean5, double* covp, double* covc, int* meta, int* ids,
                  void* stream);
int pk_map_import(void* pool, int capacity, int dtype, const int* slot, long long p_lo,
                  long long count, const double* mean5, const double* covp, const double* covc,
                  const int* meta, const int* ids, void* stream);

/* ---- K1 motion: FastSLAM.motion_update / motion_model (prkt_core_v2.py:148-208) and the
 *      quaternion heading round trip (utils.py:8-35) ------------------------------------------ */
/* noise3 != NULL: standard normals [M][3] standing for the three numpy normal() draws
 * (injected-noise parity mode).  noise3 == NULL: Philox4x32-10 counter RNG keyed by `seed`,
 * counter = (particle_offset + i, frame), Box-Muller. */
int pk_motion_update(double* pose4, long long M, const double* noise3, unsigned long long seed,
                     unsigned long long frame, long long particle_offset, double v, double w,
                     double dt, void* stream);

/* ---- K2 fused measurement update: the per-particle body of FastSLAM.cam_cb
 *      (prkt_core_v2.py:84-124): match_features_to_scan / match_one / probability_of_match /
 *      prob_position_match / closest_point / prob_color_match (:317-544), generate_measurement,
 *      measurement_jacobian, measurement_covariance, inverse, kalman_gain (:748-833, 859-877),
 *      Feature.update_mean / update_covar (:897-930), importance_factor, no_match_weight
 *      (:835-857), potential-feature promotion (:109-118), orphan counting (:740-746) ---------- */
/* obs_host: K rows of (bearing, r, g, b), HOST memory, read before the call returns.
 * Writes weight (pose4[i][3]), assoc[M][K] (reference ids: >0 full, <0 potential, 0 unseen),
 * updates the pool and aux2, and fills stats[PK_NUM_STATS] (zeroed by the call itself). */
int pk_measurement_update(double* pose4, int* aux2, const int* slot, void* pool, int capacity,
                          int dtype, long long M, const double* obs_host, int K,
                          const pk_params* params, int* assoc, unsigned long long* stats,
                          void* stream);

/* Same with the scan resident on the device (obs_dev[K][4]; e.g. written by pk_simulate_scan), so a frame
 * never visits the host.  table_ws: pk_obs_table_bytes() of device scratch. */
long long pk_obs_table_bytes(void);
int pk_measurement_update_dev(double* pose4, int* aux2, const int* slot, void* pool, int capacity,
                              int dtype, long long M, const double* obs_dev, int K,
                              const pk_params* params, int* assoc, unsigned long long* stats,
                              void* table_ws, void* stream);

/* ---- K2b spawn mode: FilterParticle.add_hypothesis for every unseen blob (id 0) of the frame, in scan
 *      order: find_nearest_reading / reading_distance_function / ray_intersect / color_distance
 *      (prkt_core_v2.py:546-651), add_new_feature + cross_readings (:653-738) or
 *      add_orphaned_reading (:740-746).  As WRITTEN this path is dead code (SURVEY.md finding F5); this
 *      entry point implements the reference with the three documented patches of SURVEY.md A.6
 *      (oracle/ref_shim.apply_spawn_patches), i.e. the behaviour the docstrings describe.  Call after
 *      pk_measurement_update (which already applied the 0.1 weight factor and the next_id bump of the
 *      unseen blobs, :95, :745-746) with the same obs_host and the assoc it wrote.  dtype must carry an
 *      orphan-slot count.  pair_gate: largest colour distance that still pairs two readings. */
int pk_spawn_update(const double* pose4, int* aux2, const int* slot, void* pool, int capacity,
                    int dtype, long long M, const double* obs_host, int K, const int* assoc,
                    double pair_gate, unsigned long long* stats, void* stream);
int pk_spawn_update_dev(const double* pose4, int* aux2, const int* slot, void* pool, int capacity,
                        int dtype, long long M, const double* obs_dev, int K, const int* assoc,
                        double pair_gate, unsigned long long* stats, void* stream);
/* Orphan readings of particles [p_lo, p_lo+count): totals[count] (readings ever stored) and
 * readings[count][slots][8] doubles in ring order (tests, FilterParticle.hypothesis_set views). */
int pk_orphans_export(const void* pool, int capacity, int dtype, const int* slot, long long p_lo,
                      long long count, int* totals, double* readings, void* stream);

/* ---- K3/K4 weight normaliser + systematic resampling plan:
 *      FastSLAM.low_variance_resample (prkt_core_v2.py:210-252) ------------------------------- */
long long pk_num_scan_blocks(long long M);
/* K3a: per-block (PK_SCAN_BLOCK particles) inclusive prefix sums of the weights, fp64, and the
 * block totals.  cumsum[M], block_sums[pk_num_scan_blocks(M)]. */
int pk_weight_scan(const double* pose4, long long M, double* cumsum, double* block_sums,
                   void* stream);
/* K3b: fold ALL block totals of ALL shards (nb_total of them, in global order; on one GPU this is
 * the block_sums array itself) into double-double block prefixes and the resampling thresholds.
 * plan: double[PK_PLAN_DOUBLES] (total, range, u0 ...); block_prefix[nb_total][2];
 * block_count[nb_total+1] (outputs emitted before each block, monotone). */
#define PK_PLAN_DOUBLES 8
int pk_resample_thresholds(const double* all_block_sums, long long nb_total, long long M_total,
                           double u01, double* plan, double* block_prefix, long long* block_count,
                           void* stream);
/* K4: for the local particles (global index particle_offset + i, scan blocks starting at
 * block_offset) compute each one's run of output slots [out_lo[i], out_lo[i]+offspring[i]) and
 * write ancestors[k - out_offset] = particle_offset + i for every output k of the local output
 * window [out_offset, out_offset + n_out).  big_runs: workspace of 4 + 3*R int64, R = the most runs of MORE than 16
 * equal ancestors the window can hold = min(M_local, n_out / 17 + 2) (a counter and R triples: ancestor, first and
 * one-past-last slot; such runs are written by a second kernel). */
int pk_resample_ancestors(const double* cumsum, long long M_local, long long particle_offset,
                          long long block_offset, const double* plan, const double* block_prefix,
                          const long long* block_count, long long M_total, long long out_offset,
                          long long n_out, long long* out_lo, int* offspring, long long* ancestors,
                          long long* big_runs, void* stream);

/* K4 and the first step of K5 in ONE kernel (what the filters call every frame): pk_resample_ancestors' outputs for the
 * window [out_offset, out_offset + n_out) -- runs of any length, no big_runs list -- plus, into gather_workspace
 * (pk_gather_workspace_bytes(M_local)), each local particle's offspring count INSIDE the window and the per-scan-block
 * exclusive count of particles that have none (their landmark block is free).  Follow with
 * pk_resample_gather_planned (single GPU: window = [0, M)) or pk_resample_gather_peer (window = this rank's slots). */
int pk_resample_plan(const double* cumsum, long long M_local, long long particle_offset,
                     long long block_offset, const double* plan, const double* block_prefix,
                     const long long* block_count, long long M_total, long long out_offset,
                     long long n_out, long long* out_lo, int* offspring, long long* ancestors,
                     void* gather_workspace, void* stream);

/* ---- K5 copy-on-resample: the deepcopy(particle) of prkt_core_v2.py:243 ---------------------- */
long long pk_gather_workspace_bytes(long long M);
/* pk_resample_gather after pk_resample_plan(window [0, M)) has filled `workspace`: free-block list (each scan block
 * adds up the dead counts before it), pose / aux / slot permutation, block copies -- three launches. */
int pk_resample_gather_planned(const long long* ancestors, long long M, const double* pose4_in,
                               double* pose4_out, const int* aux2_in, int* aux2_out,
                               const int* slot_in, int* slot_out, void* pool, int capacity, int dtype,
                               void* workspace, long long* n_copied_out, void* stream);
/* The block copies of pk_resample_gather_planned on their own: call that function with capacity = 0 (free list and
 * permutation only, two launches) and this one with the real capacity on ANOTHER stream, ordered after it by an event.
 * Nothing but the landmark pool is touched here, so the next frame's motion update (poses only) can run beside the copies;
 * whatever reads or writes the pool next must wait for `stream`.  Replaces the same deepcopy (prkt_core_v2.py:243). */
int pk_resample_copy_blocks(void* pool, int capacity, int dtype, long long M, void* workspace,
                            const long long* n_copied, void* stream);
/* Single-GPU form.  ancestors[M] ascending (int64, local indices).  Survivors keep their landmark
 * block; every extra copy of a particle is written into the block of a particle that died
 * (#copies == #dead).  pose/aux are permuted out of place.  n_copied_out (device int64) receives
 * the number of blocks copied. */
int pk_resample_gather(const long long* ancestors, const int* offspring, long long M,
                       const double* pose4_in, double* pose4_out, const int* aux2_in,
                       int* aux2_out, const int* slot_in, int* slot_out, void* pool, int capacity,
                       int dtype, void* workspace, long long* n_copied_out, void* stream);
/* ---- sharded form (particles split in contiguous index ranges over ranks, one process per GPU).
 * A migrating particle travels as one record of pk_particle_record_bytes(): a 64-byte header
 * (pose4, aux2) followed by its landmark block. */
long long pk_particle_record_bytes(int capacity, int dtype);
/* Pack the n particles whose GLOBAL indices are emit_run[0..n) (all local to this rank) into
 * out[n][record].  workspace: 3*n ints. */
int pk_pack_particles(const long long* emit_run, long long n, long long particle_offset,
                      const double* pose4, const int* aux2, const int* slot, const void* pool,
                      int capacity, int dtype, void* out, int* workspace, void* stream);
/* The rank's Ml output slots are [n_lo arrivals from lower ranks | n_loc offspring of local
 * ancestors | the rest arrivals from higher ranks] (ancestors ascend globally).  local_run[n_loc]:
 * global ancestor index of each local output; out_lo / offspring: as written by
 * pk_resample_ancestors for the local particles; recv: the arrivals' records in source-rank order.
 * Arrivals and local duplicates take blocks freed by local particles without local offspring.
 * Call AFTER the outgoing particles were packed.  total_dead_out (device int64): freed blocks.
 * workspace: pk_gather_workspace_bytes(Ml). */
int pk_resample_gather_sharded(const long long* local_run, const long long* out_lo,
                               const int* offspring, long long Ml, long long particle_offset,
                               long long n_lo, long long n_loc, const double* pose4_in,
                               double* pose4_out, const int* aux2_in, int* aux2_out,
                               const int* slot_in, int* slot_out, const void* recv, void* pool,
                               int capacity, int dtype, void* workspace,
                               long long* total_dead_out, void* stream);
/* Raw block mover (also used by the sharded path to pack / unpack migrating particles): copies n
 * landmark blocks src_base[src_slot[i]] -> dst_base[dst_slot[i]] (n = min(*n_dev, n_max) when
 * n_dev != NULL) staged through shared memory with TMA bulk copies.  n_live (nullable) limits
 * each copy to the live landmarks of the block. */
int pk_copy_blocks(const void* src_base, void* dst_base, int capacity, int dtype, const int* src_slot,
                   const int* dst_slot, const int* n_live, long long n_max, const long long* n_dev,
                   void* stream);

/* ---- peer path of the sharded filter: same arithmetic, but every count stays on the device and the
 *      two exchanges of low_variance_resample (prkt_core_v2.py:218-226 needs the global weight sum,
 *      :243 moves whole particles) are done by this library's kernels over NVLink peer memory.  Each
 *      rank owns ONE peer allocation [flags | all block totals | receive buffer] and holds device
 *      tables (uint64[n_ranks]) with the address of each region on every rank. ---------------------- */
int pk_peer_alloc(long long bytes, void** ptr_out);                   /* cudaMalloc + zero fill */
int pk_peer_free(void* ptr);
int pk_peer_export(void* ptr, unsigned char* handle_out_host);        /* PK_PEER_HANDLE_BYTES */
int pk_peer_open(const unsigned char* handle_host, void** ptr_out);   /* map a peer's allocation */
int pk_peer_close(void* ptr);
/* K3a fused with the all-gather of the block totals: block b of rank `rank` is also stored to
 * peer_sums_tab[g][rank * nb + b] for every rank g (g == rank included). */
int pk_weight_scan_publish(const double* pose4, long long M, double* cumsum, double* block_sums,
                           const unsigned long long* peer_sums_tab, int rank, int n_ranks,
                           void* stream);
/* Device-side barrier over flags in peer memory (peer_flags_tab[g] = uint64[n_ranks] on rank g).
 * `epoch` must increase by one per call, starting at 1, identically on all ranks.  A peer that does
 * not arrive within timeout_s sets PK_PEER_TIMEOUT in *status (device uint64) instead of hanging. */
int pk_peer_barrier(const unsigned long long* peer_flags_tab, int rank, int n_ranks,
                    unsigned long long epoch, double timeout_s, unsigned long long* status,
                    void* stream);
/* First half of a split-phase barrier: post this rank's flag for `epoch` (everything the stream wrote to peer memory
 * before is ordered in front of it); the matching wait is the one inside pk_resample_gather_peer. */
int pk_peer_post(const unsigned long long* peer_flags_tab, int rank, int n_ranks,
                 unsigned long long epoch, void* stream);
/* Exchange plan of this rank from block_count (K3b output over all ranks' blocks, nb_per_rank each):
 * xplan[PK_XPLAN_LONGS] device int64.  `capacity` = records a rank may RECEIVE per frame (the size
 * of its receive buffer); a rank then sends at most (n_ranks - 1) * capacity.  Exceeding it sets
 * PK_PEER_OVERFLOW in *status and voids the frame (no out-of-bounds access). */
int pk_exchange_plan(const long long* block_count, long long nb_per_rank, int n_ranks, int rank,
                     long long Ml, long long capacity, long long* xplan, unsigned long long* status,
                     void* stream);
/* Same arithmetic on the host from emitted_before[n_ranks+1] (CPU tests, debugging). */
int pk_exchange_plan_host(const long long* emitted_before_host, int n_ranks, int rank, long long Ml,
                          long long capacity, long long* xplan_host);
/* Push every offspring of a local particle whose output slot lives on another rank straight into
 * that rank's receive buffer (peer_recv_tab[g]).  send_capacity >= (n_ranks - 1) * capacity;
 * workspace: 4 * send_capacity ints. */
int pk_push_particles(const long long* xplan, const long long* out_lo, long long Ml, int rank,
                      const double* pose4, const int* aux2, const int* slot, const void* pool,
                      int capacity, int dtype, const unsigned long long* peer_recv_tab,
                      long long send_capacity, int* workspace, void* stream);
/* pk_resample_gather_sharded with the window split read from xplan; anc_window[Ml] = global ancestor
 * of each local output slot and `workspace` as left by pk_resample_plan with out_offset = rank * Ml, n_out = Ml. */
/* peer_flags_tab != NULL: the flag barrier that makes the other ranks' pushes visible (pk_peer_barrier's) runs inside
 * this call instead of in a launch of its own, in front of the part that reads the receive buffer: the free list, the
 * slots filled from local ancestors and the copies of the local duplicates run first, while the pushes are still in
 * flight.  status: PK_PEER_STATUS_WORDS device uint64.  pushes_done_event: NULL, or a cudaEvent_t recorded after this
 * rank's pk_push_particles AND pk_peer_post(epoch) when those ran on ANOTHER stream (so that they overlap the free list
 * and the slot assignment of this call, and the other ranks do not wait for this rank's local copies): `stream` waits
 * for it before any landmark block is overwritten, and this call only waits for the flags. */
int pk_resample_gather_peer(const long long* xplan, const long long* anc_window,
                            const long long* out_lo, const int* offspring, long long Ml,
                            long long particle_offset, const double* pose4_in, double* pose4_out,
                            const int* aux2_in, int* aux2_out, const int* slot_in, int* slot_out,
                            const void* recv, long long recv_capacity, void* pool, int capacity,
                            int dtype, void* workspace, long long* total_dead_out,
                            const unsigned long long* peer_flags_tab, int rank, int n_ranks,
                            unsigned long long epoch, double timeout_s, unsigned long long* status,
                            void* pushes_done_event, void* stream);
/* K3b of the peer path in ONE single-CTA kernel: the flag barrier after the fused all-gather of the block totals,
 * pk_resample_thresholds over all ranks' totals, and pk_exchange_plan (xplan, PK_PEER_OVERFLOW into status[0];
 * status: PK_PEER_STATUS_WORDS device uint64). */
int pk_resample_thresholds_peer(const double* all_block_sums, long long nb_total, long long M_total,
                                double u01, double* plan, double* block_prefix, long long* block_count,
                                const unsigned long long* peer_flags_tab, int rank, int n_ranks,
                                unsigned long long epoch, double timeout_s, long long Ml,
                                long long capacity, long long* xplan, unsigned long long* status,
                                void* stream);

/* ---- log-domain weight normaliser (PK_MODEL_LOG_WEIGHTS; north_star "log-sum-exp normalisation") ----------------
 * pk_log_weights_max: max_out[0] = max_i pose4[i][3] (warp-shuffle + block reduction, fixed order).
 * pk_log_weights_normalise: pose4[i][3] <- exp(pose4[i][3] - max_in[0]) in place (max_in: this rank's maximum, or the
 * all-reduced one of a sharded filter), out3 = sum w, sum w^2 of the LOCAL particles and max_in[0]; the caller
 * all-reduces the sums when sharded: log-sum-exp = max + log(sum w), N_eff = (sum w)^2 / sum w^2.
 * workspace: 2 * 1024 doubles. */
int pk_log_weights_max(const double* pose4, long long M, double* max_out, double* workspace, void* stream);
int pk_log_weights_normalise(double* pose4, long long M, const double* max_in, double* out3, double* workspace,
                             void* stream);

/* ---- K6 queries: FastSLAM.summary (prkt_core_v2.py:254-276); best particle is additive ------- */
/* out5[0..3] = sum x, sum y, sum sin(theta), sum cos(theta); out5[4] = M.  The caller finishes
 * (x/M, y/M, atan2) after an optional cross-shard all-reduce.  workspace: 5*1024 doubles. */
int pk_summary_partial(const double* pose4, long long M, double* out5, double* workspace,
                       void* stream);
/* best[0] = max weight, best[1] = index of its first occurrence (as double). workspace 2*1024. */
int pk_best_particle(const double* pose4, long long M, double* best2, double* workspace,
                     void* stream);

/* ---- tools around the filter -------------------------------------- */
/* Scan simulator standing in for the un-vendored viz_feature_sim node (VizScan of K Blobs, matrix.py:35-39,
 * prkt_core_v2.py:344): the K landmarks of landmarks5[N][5] (x, y, r, g, b; device) nearest the true pose, in
 * ascending distance order, bearing = wrap_pi(atan2(ly-y, lx-x) - theta) + sigma_bearing * z0, colour = truth +
 * sigma_color * z1..3.  noise4[K][4]: injected standard normals (device), or NULL for on-device counter noise
 * keyed by (seed, frame).  workspace: N doubles.  obs_out[K][4] and landmark_out[K] (nullable) stay on the device. */
int pk_simulate_scan(const double* landmarks5, int N, double x, double y, double theta, int K,
                     const double* noise4, unsigned long long seed, unsigned long long frame,
                     double sigma_bearing, double sigma_color, double* workspace, double* obs_out,
                     int* landmark_out, void* stream);
/* Accuracy analysis (the working form of analyze_slam.py:1-36, utils.py heading_error / minimize_angle):
 * out8 = sum w, sum w^2, sum (x-xt)^2, sum (y-yt)^2, sum minimize_angle(theta-tt)^2, sum x, sum y, max w over
 * the particles.  workspace: 8*512 doubles. */
int pk_accuracy(const double* pose4, long long M, double x_true, double y_true, double theta_true,
                double* out8, double* workspace, void* stream);
/* Per true landmark j (ids j+1, load_feature_list :294-299): err2_out[j] = sum over particles of the squared
 * position error of the landmark carrying that id, count_out[j] = particles that hold it. */
int pk_map_error(const void* pool, int capacity, int dtype, const int* slot, const int* aux2,
                 long long M, const double* truth5, int N, double* err2_out,
                 unsigned long long* count_out, void* stream);

/* ---- probes: the per-landmark math one triple per thread (back the scalar helper methods of
 *      FilterParticle -- probability_of_match prkt_core_v2.py:383-455, the EKF pieces :748-930 --
 *      and pin the device arithmetic against the reference's known-answer vectors) ------------- */
/* pose3[n][3] (x,y,heading), blob4[n][4] (bearing,r,g,b), dir2[n][2] = unit((cos b, sin b, 0)),
 * landmark mean5[n][5], covp[n][4], covc[n][9]  ->  out[n] = probability_of_match */
int pk_probe_likelihood(const double* pose3, const double* blob4, const double* dir2,
                        const double* mean5, const double* covp, const double* covc, long long n,
                        const pk_params* params, double* out, void* stream);
/* One EKF update per thread: pose2[n][2], blob4[n][4], landmark (mean5, covp, covc, meta nullable)
 * -> updated landmark and the weight factor (importance_factor, or no_match_weight if potential). */
int pk_probe_ekf(const double* pose2, const double* blob4, const double* mean5, const double* covp,
                 const double* covc, const int* meta, long long n, const pk_params* params,
                 double* mean5_out, double* covp_out, double* covc_out, double* factor_out,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PARAKEET_B200_H */
