// K3/K4/K5/K6 -- weight normaliser, systematic (low-variance) resampling plan, copy-on-resample
// gather and the pose summary.
//
// Replaces FastSLAM.low_variance_resample (reference prkt_core_v2.py:210-252) and
// FastSLAM.summary (:254-276).  The reference sweeps the particle list once with a running
// `step` (:238-250); that is `ancestor[k] = min{ i : C_i >= u0 + k*r }` with C the prefix sum of
// the weights, r = sum/M, u0 = random()*r (SURVEY.md finding F6).  Here:
//
//   K3a  one warp per block of PK_SCAN_BLOCK particles: fp64 inclusive prefix sums that are
//        monotone by construction (each lane folds 32 consecutive weights left to right, lane
//        bases are a left fold of the lane totals);
//   K3b  one cluster of 8 CTAs folds the block totals of ALL shards, in global block order and with a fixed
//        tree, into double-double block prefixes -> total, r, u0 and the number of outputs emitted
//        before every block.  Nothing depends on how many GPUs the particles are spread over;
//   K4   one thread per particle: its run of output slots is [N(C_{i-1}), N(C_i)) with
//        N(C) = #{k : u0 + k*r <= C}, evaluated in double-double against the block prefix.  No
//        search, no atomics on the common path; runs longer than 16 go to a fill kernel;
//   K5   survivors keep their landmark block; each additional copy goes into the block of a
//        particle that died (an exclusive scan of the dead flags pairs them up), moved by TMA
//        bulk copies global -> shared -> global.  This is the deepcopy of :243.
#include "pk_common.cuh"
#include "pk_filter_math.cuh"

namespace pk {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kScanWarps = 4;
constexpr int kMinGroupBlocks = 8;  // scan blocks per group in K3b (doubles while > 1024 groups)
constexpr int kMaxGroupBlocks = 32;
constexpr int kMaxScanGroups = 1024;

__device__ __forceinline__ int padded(int e) { return e + (e >> 5); }

// ---------------------------------------------------------------------------------------------
// K3a
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanWarps * 32)
weight_scan_kernel(const double* __restrict__ pose4, long long M, double* __restrict__ cumsum,
                   double* __restrict__ block_sums, long long nb, double* const* __restrict__ peer_sums, int me,
                   int n_ranks) {
    __shared__ double sm[kScanWarps][PK_SCAN_BLOCK + 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long blk = (long long)blockIdx.x * kScanWarps + warp;
    if (blk >= nb) return;
    const long long base = blk * PK_SCAN_BLOCK;
    const int n = (int)min((long long)PK_SCAN_BLOCK, M - base);
    double* s = sm[warp];
#pragma unroll 4
    for (int it = 0; it < 32; ++it) {
        const int e = it * 32 + lane;
        s[padded(e)] = (e < n) ? pose4[4 * (base + e) + 3] : 0.0;
    }
    __syncwarp();
    double run = 0.0;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        const int e = padded(32 * lane + j);
        run = __dadd_rn(run, s[e]);   // left fold, as sum_ += weight (:218-220)
        s[e] = run;
    }
    const double total = run;
    double b = 0.0;
    for (int l = 0; l < 31; ++l) {
        const double t = __shfl_sync(kFullMask, total, l);
        if (lane > l) b = __dadd_rn(b, t);
    }
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        const int e = padded(32 * lane + j);
        s[e] = __dadd_rn(b, s[e]);
    }
    __syncwarp();
#pragma unroll 4
    for (int it = 0; it < 32; ++it) {
        const int e = it * 32 + lane;
        if (e < n) cumsum[base + e] = s[padded(e)];
    }
    const double block_total = __dadd_rn(b, total);
    if (lane == 31) block_sums[blk] = block_total;
    if (peer_sums != nullptr) {
        // fused all-gather: lane g stores this block's total straight into rank g's copy of the global
        // block-total array (peer memory over NVLink; rank g == me is the local copy)
        const double t31 = __shfl_sync(kFullMask, block_total, 31);
        if (lane < n_ranks) peer_sums[lane][(long long)me * nb + blk] = t31;
    }
}

// number of outputs k in [0, M) with u0 + k*r <= P + c   (P double-double block prefix)
__device__ __forceinline__ long long count_le(dd P, double c, double u0, double r, long long M) {
    if (!(r > 0.0)) return M;  // all-zero weights: step == 0 <= 0, particle 0 is emitted M times (:239)
    const dd t = dd_add_d(P, c);
    const double x = (t.hi - u0) + t.lo;
    double kq = floor(x / r);
    if (!(kq >= -1.0)) kq = -1.0;
    if (kq > (double)(M - 1)) kq = (double)(M - 1);
    long long k = (long long)kq;
    auto pred = [&](long long kk) { return (fma((double)kk, r, u0) - t.hi) <= t.lo; };
    for (int it = 0; it < 64 && k + 1 < M && pred(k + 1); ++it) ++k;
    for (int it = 0; it < 64 && k >= 0 && !pred(k); ++it) --k;
    return k + 1;
}

// ---- sharded path -------------------------------------------------------------------------------
// A migrating particle travels as one RECORD: a 64-byte header (pose4: 32 B, aux2: 8 B, padding) followed
// by its landmark block.  kHeaderBytes keeps the block 64-byte aligned inside the exchange buffer.
constexpr int kHeaderBytes = 64;

// ---- device-resident exchange plan (peer path) ------------------------------------------------------
// Everything the exchange needs follows from E[g] = number of output slots whose ancestor lives on a
// rank < g (E[g] = block_count[g * nb], which every rank computes identically in K3b): rank g's
// offspring occupy the global output slots [E[g], E[g+1]) and output slot k belongs to rank k / Ml.
// The same function runs on the host (pk_exchange_plan_host; CPU tests compare it with the Python
// plan_exchange) and in a one-thread kernel, so no count ever crosses PCIe.
enum {
    XP_EMIT_LO = 0,    // E[me]
    XP_EMIT_N = 1,     // E[me+1] - E[me]: outputs descending from my particles
    XP_N_LO = 2,       // my output slots filled from lower ranks
    XP_N_LOC = 3,      // ... from my own particles
    XP_N_HI = 4,       // ... from higher ranks
    XP_N_BELOW = 5,    // my offspring that live on lower ranks
    XP_N_ABOVE = 6,    // ... on higher ranks
    XP_ABOVE_START = 7,  // first global output slot of the N_ABOVE run
    XP_N_SEND = 8,     // N_BELOW + N_ABOVE (0 when XP_OVERFLOW)
    XP_N_IN = 9,       // N_LO + N_HI        (0 when XP_OVERFLOW)
    XP_OVERFLOW = 10,  // some rank would receive more than the exchange capacity
    XP_N_RANKS = 11,
    XP_ML = 12,
    XP_RANK_LO = 16,   // [PK_MAX_RANKS] n_lo of every rank
    XP_RANK_LOC = 16 + PK_MAX_RANKS,  // [PK_MAX_RANKS] n_loc of every rank
};
static_assert(PK_XPLAN_LONGS >= 16 + 2 * PK_MAX_RANKS, "exchange plan size");

__host__ __device__ inline long long xp_clamp(long long v, long long lo, long long hi) { return v < lo ? lo : (v > hi ? hi : v); }

__host__ __device__ inline void make_exchange_plan(const long long* E, int G, long long Ml, int me, long long cap,
                                                   long long* xp) {
    for (int i = 0; i < PK_XPLAN_LONGS; ++i) xp[i] = 0;
    bool overflow = false;
    for (int g = 0; g < G; ++g) {
        const long long w0 = (long long)g * Ml, w1 = w0 + Ml;
        const long long n_lo = xp_clamp(E[g] - w0, 0, Ml);             // window slots below E[g]
        const long long loc_hi = xp_clamp(E[g + 1], w0, w1), loc_lo = xp_clamp(E[g], w0, w1);
        const long long n_loc = loc_hi - loc_lo;
        xp[XP_RANK_LO + g] = n_lo;
        xp[XP_RANK_LOC + g] = n_loc;
        // capacity bounds what a rank RECEIVES; what it sends is then at most (G-1) * cap (the send
        // lists are sized for that: one particle may own every output slot of the filter)
        if (Ml - n_loc > cap) overflow = true;
        if (g == me) {
            xp[XP_EMIT_LO] = E[g];
            xp[XP_EMIT_N] = E[g + 1] - E[g];
            xp[XP_N_LO] = n_lo;
            xp[XP_N_LOC] = n_loc;
            xp[XP_N_HI] = Ml - n_lo - n_loc;
            const long long below_end = xp_clamp(w0, E[g], E[g + 1]);    // offspring slots < my window
            const long long above_start = xp_clamp(w1, E[g], E[g + 1]);  // offspring slots >= window end
            xp[XP_N_BELOW] = below_end - E[g];
            xp[XP_N_ABOVE] = E[g + 1] - above_start;
            xp[XP_ABOVE_START] = above_start;
        }
    }
    xp[XP_OVERFLOW] = overflow ? 1 : 0;
    xp[XP_N_SEND] = overflow ? 0 : xp[XP_N_BELOW] + xp[XP_N_ABOVE];
    xp[XP_N_IN] = overflow ? 0 : xp[XP_N_LO] + xp[XP_N_HI];
    xp[XP_N_RANKS] = G;
    xp[XP_ML] = Ml;
}

// Cross-rank barrier on flags in peer memory: thread g tells rank g "I am at `epoch`" and waits until
// rank g said the same.  Everything this rank wrote to peer memory earlier on the stream is ordered
// before the flag (fence + release); everything the peers wrote before their flag is visible after
// the acquire.  A rank that never shows up trips the time-out instead of hanging the GPU.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Body of the cross-rank flag barrier for threads g < G of ONE CTA (callers follow it with __syncthreads()):
// thread g tells rank g "I am at `epoch`" and waits until rank g said the same.
struct PeerSync {
    unsigned long long* const* flags;   // device table: flags base of every rank (NULL: no barrier)
    int me, G;
    unsigned long long epoch, timeout_ns;
    unsigned long long* status;         // [0] sticky PK_PEER_* bits, [1 + which] nanoseconds spent waiting, [3] barriers
    int which;                          // 0: first barrier of the frame (block totals), 1: second (pushes landed)
    bool posted;                        // the rank's flag was posted earlier (pk_peer_post): wait only
};
// accounting of one CTA's wait (thread 0, after the __syncthreads() that follows peer_sync_thread)
__device__ __forceinline__ void peer_sync_account(const PeerSync& ps, unsigned long long t_entry) {
    atomicAdd(ps.status + 1 + ps.which, global_timer_ns() - t_entry);
    if (ps.which == 0) atomicAdd(ps.status + 3, 1ull);
}
__device__ __forceinline__ void peer_sync_thread(const PeerSync& ps, int g, bool post) {
    if (g >= ps.G) return;
    if (post) {
        __threadfence_system();
        st_release_sys(ps.flags[g] + ps.me, ps.epoch);
    }
    const unsigned long long* mine = ps.flags[ps.me] + g;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(mine) < ps.epoch) {
        if (global_timer_ns() - t0 > ps.timeout_ns) {
            atomicOr(ps.status, (unsigned long long)PK_PEER_TIMEOUT);
            break;
        }
        __nanosleep(64);
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// K3b  (ONE thread-block cluster: kThrCluster CTAs of 1024 threads, distributed shared memory)
//
// Block totals s_b (b in global block order) -> double-double exclusive block prefixes P_b, total,
// r, u0 and the number of outputs emitted up to the end of every block.  The fold tree is a fixed
// function of the TOTAL number of blocks (never of the launch geometry or of how the blocks are
// spread over ranks; every rank of a sharded filter runs this kernel on the same array):
//   group  = GB consecutive blocks (GB = 8, doubling while there would be more than 1024 groups):
//            Kogge-Stone inclusive scan over the GB lanes of the group        -> L_b, T_g
//   super  = 32 consecutive groups: Kogge-Stone scan over the lanes of a warp -> X_g, S_w
//   top    = Kogge-Stone scan of the <= 32 super totals (one warp)            -> Y_w, total
//   P_b    = (Y_w + X_g) + L_b
// A single SM's fp64 pipe made the former single-CTA form cost 15 us at 1024 blocks and 60-80 us
// at 8192 (8 ranks); spread over 8 SMs with 5-step scans it is a few microseconds at any size.
// ---------------------------------------------------------------------------------------------
constexpr int kThrCluster = 8;
constexpr int kThrGroupsPerCta = kMaxScanGroups / kThrCluster;   // 128 groups = 4 supers per CTA
constexpr int kThrMaxRounds = kMaxGroupBlocks / 8;               // blocks of a CTA / 1024 threads
static_assert(kThrGroupsPerCta * kMinGroupBlocks == 1024, "one block per thread and round");

__device__ __forceinline__ dd shfl_up_dd(dd v, int o) {
    return dd{__shfl_up_sync(kFullMask, v.hi, o), __shfl_up_sync(kFullMask, v.lo, o)};
}
// inclusive scan over aligned segments of `width` lanes (power of two <= 32): element = earlier + later
__device__ __forceinline__ dd seg_scan_dd(dd v, int lane, int width) {
    const int pos = lane & (width - 1);
    for (int o = 1; o < width; o <<= 1) {
        const dd up = shfl_up_dd(v, o);
        if (pos >= o) v = dd_add(up, v);
    }
    return v;
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a shared-memory object of this kernel) in CTA `rank` of the cluster
template <typename T>
__device__ __forceinline__ unsigned dsmem_addr(T* p, unsigned rank) {
    const unsigned local = (unsigned)__cvta_generic_to_shared(p);
    unsigned remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    return remote;
}
__device__ __forceinline__ void dsmem_store(unsigned addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void dsmem_store(unsigned addr, long long v) {
    asm volatile("st.shared::cluster.s64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

__global__ void __cluster_dims__(kThrCluster, 1, 1) __launch_bounds__(1024)
thresholds_kernel(const double* __restrict__ sums, long long nb, long long M_total, double u01,
                  double* __restrict__ plan, double* __restrict__ block_prefix, long long* block_count, int GB,
                  PeerSync ps, long long Ml, long long xcap, long long* __restrict__ xplan) {
    __shared__ double T_hi[kThrGroupsPerCta], T_lo[kThrGroupsPerCta];  // group totals, then group prefixes
    __shared__ double S_hi[32], S_lo[32];                              // super totals of the cluster, then prefixes
    __shared__ long long w_max[32];
    __shared__ long long c_max[kThrCluster];
    __shared__ double s_r, s_u0;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned cta = cluster_ctarank();
    if (ps.flags != nullptr) {
        // sharded filter, peer exchange: the block totals of all ranks were stored into `sums` by their scan kernels
        // (fused all-gather); CTA 0 posts this rank's flag, every CTA waits for every rank's flag before reading them
        const unsigned long long t_entry = global_timer_ns();
        peer_sync_thread(ps, t, cta == 0);
        __syncthreads();
        if (cta == 0 && t == 0) peer_sync_account(ps, t_entry);
    }
    const int rounds = GB / kMinGroupBlocks;                       // 1024 blocks of this CTA per round
    const long long cta_block0 = (long long)cta * kThrGroupsPerCta * GB;
    // phase A: one thread per block; inclusive scan inside each group of GB lanes
    double sv[kThrMaxRounds];
    dd L[kThrMaxRounds];
#pragma unroll
    for (int k = 0; k < kThrMaxRounds; ++k) {
        sv[k] = 0.0;
        L[k] = dd{0.0, 0.0};
        if (k < rounds) {
            const int j = k * 1024 + t;
            const long long b = cta_block0 + j;
            sv[k] = (b < nb) ? sums[b] : 0.0;
            const dd inc = seg_scan_dd(dd{sv[k], 0.0}, lane, GB);
            const dd up = shfl_up_dd(inc, 1);
            if ((lane & (GB - 1)) != 0) L[k] = up;
            if ((lane & (GB - 1)) == GB - 1) {
                T_hi[j / GB] = inc.hi;
                T_lo[j / GB] = inc.lo;
            }
        }
    }
    cluster_sync_all();   // (also: every CTA of the cluster is running before its shared memory is written remotely)
    // phase B1: warp w < 4 scans the 32 groups of super cta * 4 + w; its total goes to every CTA of the cluster
    if (warp < kThrGroupsPerCta / 32) {
        const int g = warp * 32 + lane;
        const dd inc = seg_scan_dd(dd{T_hi[g], T_lo[g]}, lane, 32);
        dd ex = shfl_up_dd(inc, 1);
        if (lane == 0) ex = dd{0.0, 0.0};
        T_hi[g] = ex.hi;
        T_lo[g] = ex.lo;
        const double th = __shfl_sync(kFullMask, inc.hi, 31), tl = __shfl_sync(kFullMask, inc.lo, 31);
        if (lane < kThrCluster) {
            const int w = (int)cta * (kThrGroupsPerCta / 32) + warp;
            dsmem_store(dsmem_addr(&S_hi[w], (unsigned)lane), th);
            dsmem_store(dsmem_addr(&S_lo[w], (unsigned)lane), tl);
        }
    }
    cluster_sync_all();
    // phase B2: every CTA scans the 32 super totals itself (same operations, same bits)
    if (warp == 0) {
        const dd inc = seg_scan_dd(dd{S_hi[lane], S_lo[lane]}, lane, 32);
        dd ex = shfl_up_dd(inc, 1);
        if (lane == 0) ex = dd{0.0, 0.0};
        S_hi[lane] = ex.hi;
        S_lo[lane] = ex.lo;
        if (lane == 31) {
            const double total = inc.hi + inc.lo;
            const double r = total / (double)M_total;  // range_ = sum_/float(len(particles)) :225
            const double u0 = u01 * r;                 // step = random()*range_             :226
            s_r = r;
            s_u0 = u0;
            if (cta == 0) {
                plan[0] = total;
                plan[1] = r;
                plan[2] = u0;
                plan[3] = inc.hi;
                plan[4] = inc.lo;
                plan[5] = (double)M_total;
                plan[6] = u01;
                plan[7] = 0.0;
            }
        }
    }
    __syncthreads();
    if (t < kThrGroupsPerCta) {
        const int w = (int)cta * (kThrGroupsPerCta / 32) + (t >> 5);
        const dd G = dd_add(dd{S_hi[w], S_lo[w]}, dd{T_hi[t], T_lo[t]});
        T_hi[t] = G.hi;
        T_lo[t] = G.lo;
    }
    __syncthreads();
    // phase C: global block prefixes, the emitted-output count at the end of every block, running maximum
    const double r = s_r, u0 = s_u0;
    long long e_inc[kThrMaxRounds];
    long long carry = 0;
#pragma unroll
    for (int k = 0; k < kThrMaxRounds; ++k) {
        e_inc[k] = 0;
        if (k < rounds) {
            const int j = k * 1024 + t;
            const long long b = cta_block0 + j;
            long long e = 0;
            if (b < nb) {
                const dd P = dd_add(dd{T_hi[j / GB], T_lo[j / GB]}, L[k]);
                block_prefix[2 * b] = P.hi;
                block_prefix[2 * b + 1] = P.lo;
                e = count_le(P, sv[k], u0, r, M_total);
            }
            for (int o = 1; o < 32; o <<= 1) {
                const long long up = __shfl_up_sync(kFullMask, e, o);
                if (lane >= o) e = max(e, up);
            }
            if (lane == 31) w_max[warp] = e;
            __syncthreads();
            long long base = carry, all = carry;
            for (int w = 0; w < 32; ++w) {
                const long long v = w_max[w];
                if (w < warp) base = max(base, v);
                all = max(all, v);
            }
            e_inc[k] = max(e, base);
            carry = all;
            __syncthreads();
        }
    }
    // running maximum across the CTAs of the cluster
    if (t < kThrCluster) dsmem_store(dsmem_addr(&c_max[cta], (unsigned)t), carry);
    cluster_sync_all();
    long long cbase = 0;
    for (unsigned c = 0; c < cta; ++c) cbase = max(cbase, c_max[c]);
#pragma unroll
    for (int k = 0; k < kThrMaxRounds; ++k) {
        if (k < rounds) {
            const long long b = cta_block0 + k * 1024 + t;
            if (b < nb) block_count[b + 1] = (b == nb - 1) ? M_total : max(e_inc[k], cbase);  // every output slot is assigned
        }
    }
    if (cta == 0 && t == 0) block_count[0] = 0;
    if (xplan != nullptr) {
        // the exchange plan follows from the emitted-output counts at the rank boundaries (one thread)
        __threadfence();
        cluster_sync_all();
        if (cta == 0 && t == 0) {
            const long long nbr = nb / ps.G;
            long long E[PK_MAX_RANKS + 1];
            for (int g = 0; g <= ps.G; ++g) E[g] = __ldcg(block_count + (long long)g * nbr);
            long long xp[PK_XPLAN_LONGS];
            make_exchange_plan(E, ps.G, Ml, ps.me, xcap, xp);
            for (int i = 0; i < PK_XPLAN_LONGS; ++i) xplan[i] = xp[i];
            if (xp[XP_OVERFLOW]) atomicOr(ps.status, (unsigned long long)PK_PEER_OVERFLOW);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4
// ---------------------------------------------------------------------------------------------
constexpr int kInlineRun = 16;

__global__ void __launch_bounds__(256)
ancestors_kernel(const double* __restrict__ cumsum, long long M_local, long long particle_offset, long long block_offset,
                 const double* __restrict__ plan, const double* __restrict__ block_prefix,
                 const long long* __restrict__ block_count, long long M_total, long long out_offset, long long n_out,
                 long long* __restrict__ out_lo, int* __restrict__ offspring, long long* __restrict__ ancestors,
                 long long* __restrict__ big_runs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M_local) return;
    const double r = plan[1], u0 = plan[2];
    const long long b = block_offset + i / PK_SCAN_BLOCK;
    const int within = (int)(i % PK_SCAN_BLOCK);
    const dd P{block_prefix[2 * b], block_prefix[2 * b + 1]};
    const long long c_lo = block_count[b], c_hi = block_count[b + 1];
    const long long gi = particle_offset + i;
    const bool last = (within == PK_SCAN_BLOCK - 1) || (gi == M_total - 1);
    long long hi = last ? c_hi : min(max(count_le(P, cumsum[i], u0, r, M_total), c_lo), c_hi);
    long long lo = (within == 0) ? c_lo : min(max(count_le(P, cumsum[i - 1], u0, r, M_total), c_lo), c_hi);
    if (hi < lo) hi = lo;
    out_lo[i] = lo;
    offspring[i] = (int)(hi - lo);
    const long long klo = max(lo, out_offset), khi = min(hi, out_offset + n_out);
    if (khi - klo <= kInlineRun) {
        for (long long k = klo; k < khi; ++k) ancestors[k - out_offset] = gi;
    } else {
        const unsigned long long idx = atomicAdd(reinterpret_cast<unsigned long long*>(big_runs), 1ull);
        big_runs[4 + 3 * idx + 0] = gi;
        big_runs[4 + 3 * idx + 1] = klo - out_offset;
        big_runs[4 + 3 * idx + 2] = khi - out_offset;
    }
}

__global__ void __launch_bounds__(256)
fill_runs_kernel(const long long* __restrict__ big_runs, long long* __restrict__ ancestors) {
    const long long n = big_runs[0];
    for (long long run = blockIdx.x; run < n; run += gridDim.x) {
        const long long gi = big_runs[4 + 3 * run], klo = big_runs[4 + 3 * run + 1], khi = big_runs[4 + 3 * run + 2];
        for (long long k = klo + threadIdx.x; k < khi; k += blockDim.x) ancestors[k] = gi;
    }
}

// ---------------------------------------------------------------------------------------------
// K5: pairing of extra copies with dead particles' blocks
// ---------------------------------------------------------------------------------------------
// G1: one warp per 1024 particles: exclusive count of dead particles inside the block + block total
__global__ void __launch_bounds__(kScanWarps * 32)
dead_scan_kernel(const int* __restrict__ offspring, long long M, int* __restrict__ dead_excl, int* __restrict__ block_dead,
                 long long nb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long blk = (long long)blockIdx.x * kScanWarps + warp;
    if (blk >= nb) return;
    const long long base = blk * PK_SCAN_BLOCK;
    int run = 0;
    for (int it = 0; it < 32; ++it) {
        const long long i = base + it * 32 + lane;
        const int dead = (i < M && offspring[i] == 0) ? 1 : 0;
        const unsigned bal = __ballot_sync(kFullMask, dead);
        if (i < M) dead_excl[i] = run + __popc(bal & lanemask_lt());
        run += __popc(bal);
    }
    if (lane == 0) block_dead[blk] = run;
}

// G2: exclusive scan of the block totals (single CTA)
__global__ void __launch_bounds__(1024)
block_offsets_kernel(const int* __restrict__ block_dead, long long nb, int* __restrict__ block_off,
                     long long* __restrict__ total_out) {
    __shared__ int warp_tot[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long per = (nb + 1023) / 1024;
    const long long b0 = t * per, b1 = min(nb, b0 + per);
    int s = 0;
    for (long long b = b0; b < b1; ++b) s += block_dead[b];
    // exclusive scan of the 1024 per-thread sums: warp scan, then a scan of the 32 warp totals
    int incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int wt = warp_tot[lane];
        int wi = wt;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFullMask, wi, o);
            if (lane >= o) wi += v;
        }
        warp_tot[lane] = wi - wt;  // exclusive
        if (lane == 31 && total_out) *total_out = wi;
    }
    __syncthreads();
    int run = warp_tot[warp] + incl - s;
    for (long long b = b0; b < b1; ++b) {
        block_off[b] = run;
        run += block_dead[b];
    }
}

// G3: dead particle i donates its block: free_list[rank of i among the dead] = slot[i]
__global__ void __launch_bounds__(256)
free_list_kernel(const int* __restrict__ offspring, const int* __restrict__ slot_in, long long M,
                 int* __restrict__ dead_excl, const int* __restrict__ block_off, int* __restrict__ free_list) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int g = dead_excl[i] + block_off[i / PK_SCAN_BLOCK];
    dead_excl[i] = g;  // now global
    if (offspring[i] == 0) free_list[g] = slot_in[i];
}

// K4 + G1 in one kernel (one CTA of 1024 threads per scan block): every particle's run of output slots, the
// ancestors of the output window (runs of any length: long ones are filled by the whole CTA), and the exclusive count
// of DEAD particles -- particles without an output inside the window, whose landmark block is therefore free -- inside
// the block.  Replaces ancestors_kernel + fill_runs_kernel + (offspring_window_kernel +) dead_scan_kernel and the
// big-run list in global memory.
__global__ void __launch_bounds__(PK_SCAN_BLOCK)
resample_plan_kernel(const double* __restrict__ cumsum, long long M_local, long long particle_offset, long long block_offset,
                     const double* __restrict__ plan, const double* __restrict__ block_prefix,
                     const long long* __restrict__ block_count, long long M_total, long long out_offset, long long n_out,
                     long long* __restrict__ out_lo, int* __restrict__ offspring, long long* __restrict__ ancestors,
                     int* __restrict__ offspring_window, int* __restrict__ dead_excl, int* __restrict__ block_dead) {
    __shared__ long long s_klo[PK_SCAN_BLOCK], s_khi[PK_SCAN_BLOCK];
    __shared__ int s_who[PK_SCAN_BLOCK];
    __shared__ int s_nbig;
    __shared__ int s_warp[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long blk = blockIdx.x;
    const long long i = blk * PK_SCAN_BLOCK + t;
    __shared__ long long s_edge[32];
    if (t == 0) s_nbig = 0;
    const bool in = i < M_local;
    long long klo = 0, khi = 0;
    // N(C_i) once per particle: the start of particle i's run is the end of particle i-1's (shuffle; shared memory
    // across warp boundaries), the block's first run starts at the block's emitted-output count
    long long c_lo = 0, c_hi = 0, raw = 0;
    bool last = false;
    if (in) {
        const double r = plan[1], u0 = plan[2];
        const long long b = block_offset + blk;
        const dd P{block_prefix[2 * b], block_prefix[2 * b + 1]};
        c_lo = block_count[b];
        c_hi = block_count[b + 1];
        last = (t == PK_SCAN_BLOCK - 1) || (particle_offset + i == M_total - 1);
        raw = min(max(count_le(P, cumsum[i], u0, r, M_total), c_lo), c_hi);
    }
    if (lane == 31) s_edge[warp] = raw;
    __syncthreads();
    long long prev = __shfl_up_sync(kFullMask, raw, 1);
    if (lane == 0 && warp > 0) prev = s_edge[warp - 1];
    if (in) {
        const long long gi = particle_offset + i;
        long long hi = last ? c_hi : raw;
        const long long lo = (t == 0) ? c_lo : prev;
        if (hi < lo) hi = lo;
        out_lo[i] = lo;
        offspring[i] = (int)(hi - lo);
        klo = max(lo, out_offset);
        khi = min(hi, out_offset + n_out);
        if (khi < klo) khi = klo;
        offspring_window[i] = (int)(khi - klo);
        if (khi - klo <= kInlineRun) {
            for (long long k = klo; k < khi; ++k) ancestors[k - out_offset] = gi;
        } else {
            const int idx = atomicAdd(&s_nbig, 1);
            s_who[idx] = t;
            s_klo[idx] = klo - out_offset;
            s_khi[idx] = khi - out_offset;
        }
    }
    // dead particles of this block: exclusive count
    const bool dead = in && (khi == klo);
    const unsigned bal = __ballot_sync(kFullMask, dead);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        const int v = s_warp[lane];
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += u;
        }
        s_warp[lane] = incl - v;
        if (lane == 31) block_dead[blk] = incl;
    }
    __syncthreads();
    if (in) dead_excl[i] = s_warp[warp] + __popc(bal & lanemask_lt());
    // long runs (a particle with more than kInlineRun offspring in the window): the whole CTA writes them
    const int nbig = s_nbig;
    for (int q = 0; q < nbig; ++q) {
        const long long gi = particle_offset + blk * PK_SCAN_BLOCK + s_who[q];
        const long long a = s_klo[q], e = s_khi[q];
        for (long long k = a + t; k < e; k += PK_SCAN_BLOCK) ancestors[k] = gi;
    }
}

// G2 + G3 in one kernel (one CTA per scan block): the block's offset among the dead is the sum of the block totals
// before it (every CTA adds them up itself: nb ints out of L2, instead of a single-CTA scan kernel in between), and
// dead particle i donates its block: free_list[rank of i among the dead] = slot[i].
__global__ void __launch_bounds__(PK_SCAN_BLOCK)
free_list_fused_kernel(const int* __restrict__ offspring_window, const int* __restrict__ slot_in, long long M,
                       int* __restrict__ dead_excl, const int* __restrict__ block_dead, long long nb,
                       int* __restrict__ free_list, long long* __restrict__ total_out) {
    __shared__ int s_part[32];
    __shared__ int s_off;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long blk = blockIdx.x;
    int s = 0;
    for (long long b = t; b < blk; b += PK_SCAN_BLOCK) s += block_dead[b];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
    if (lane == 0) s_part[warp] = s;
    __syncthreads();
    if (warp == 0) {
        int v = s_part[lane];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
        if (lane == 0) {
            s_off = v;
            if (blk == nb - 1 && total_out) *total_out = (long long)v + block_dead[blk];
        }
    }
    __syncthreads();
    const long long i = blk * PK_SCAN_BLOCK + t;
    if (i >= M) return;
    const int g = dead_excl[i] + s_off;
    dead_excl[i] = g;  // now global
    if (offspring_window[i] == 0) free_list[g] = slot_in[i];
}

// G4: per output slot k: permute pose/aux, keep or allocate a landmark block
__global__ void __launch_bounds__(256)
assign_kernel(const long long* __restrict__ ancestors, long long M, const double* __restrict__ pose_in,
              double* __restrict__ pose_out, const int* __restrict__ aux_in, int* __restrict__ aux_out,
              const int* __restrict__ slot_in, int* __restrict__ slot_out, const int* __restrict__ dead_excl,
              const int* __restrict__ free_list, int* __restrict__ copy_src, int* __restrict__ copy_dst,
              int* __restrict__ copy_nlive) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const long long a = ancestors[k];
    const double2* src = reinterpret_cast<const double2*>(pose_in + 4 * a);
    double2* dst = reinterpret_cast<double2*>(pose_out + 4 * k);
    dst[0] = src[0];
    dst[1] = src[1];
    const int2 ax = reinterpret_cast<const int2*>(aux_in)[a];
    reinterpret_cast<int2*>(aux_out)[k] = ax;
    const bool first = (k == 0) || (ancestors[k - 1] != a);
    if (first) {
        slot_out[k] = slot_in[a];
    } else {
        // outputs before k: k; of those, one per live ancestor <= a keeps its block
        const long long alive_before = a - dead_excl[a];
        const long long nidx = k - alive_before - 1;
        const int d = free_list[nidx];
        slot_out[k] = d;
        copy_src[nidx] = slot_in[a];
        copy_dst[nidx] = d;
        copy_nlive[nidx] = ax.x;
    }
}

__global__ void exchange_plan_kernel(const long long* __restrict__ block_count, long long nb, int G, int me, long long Ml,
                                     long long cap, long long* __restrict__ xplan, unsigned long long* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long E[PK_MAX_RANKS + 1];
    for (int g = 0; g <= G; ++g) E[g] = block_count[(long long)g * nb];
    long long xp[PK_XPLAN_LONGS];
    make_exchange_plan(E, G, Ml, me, cap, xp);
    for (int i = 0; i < PK_XPLAN_LONGS; ++i) xplan[i] = xp[i];
    if (xp[XP_OVERFLOW]) atomicOr(status, (unsigned long long)PK_PEER_OVERFLOW);
}

__global__ void __launch_bounds__(PK_MAX_RANKS)
peer_barrier_kernel(unsigned long long* const* __restrict__ peer_flags, int me, int G, unsigned long long epoch,
                    unsigned long long timeout_ns, unsigned long long* __restrict__ status) {
    peer_sync_thread(PeerSync{peer_flags, me, G, epoch, timeout_ns, status, 0, false}, (int)threadIdx.x, true);
}

// First half of a split-phase barrier: post this rank's flag only (the wait sits in a later kernel, possibly of
// another stream).
__global__ void __launch_bounds__(PK_MAX_RANKS)
peer_post_kernel(unsigned long long* const* __restrict__ peer_flags, int me, int G, unsigned long long epoch) {
    const int g = (int)threadIdx.x;
    if (g >= G) return;
    __threadfence_system();
    st_release_sys(peer_flags[g] + me, epoch);
}

// One thread per migrating particle.  Send item j is global output slot k (the N_BELOW run starts at
// E[me], the N_ABOVE run at ABOVE_START); its ancestor is the last local particle whose run starts at
// or before k (runs tile [E[me], E[me+1]) in index order).  The header goes straight into the
// destination rank's receive buffer; the landmark block follows through copy_blocks_kernel.
__global__ void __launch_bounds__(256)
push_headers_kernel(const long long* __restrict__ xplan, const long long* __restrict__ out_lo, long long Ml, int me,
                    const double* __restrict__ pose4, const int* __restrict__ aux2, const int* __restrict__ slot,
                    const unsigned long long* __restrict__ peer_recv, long long stride, long long cap,
                    int* __restrict__ src_slot, int* __restrict__ dst_idx, int* __restrict__ dst_rank,
                    int* __restrict__ nlive) {
    const long long n_send = min(xplan[XP_N_SEND], cap);
    const long long n_below = xplan[XP_N_BELOW];
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_send; j += (long long)gridDim.x * blockDim.x) {
        const long long k = (j < n_below) ? xplan[XP_EMIT_LO] + j : xplan[XP_ABOVE_START] + (j - n_below);
        // upper_bound(out_lo, k) - 1
        long long lo = 0, hi = Ml;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (out_lo[mid] <= k) lo = mid + 1; else hi = mid;
        }
        const long long a = lo - 1;
        const int h = (int)(k / Ml);
        const long long k_local = k - (long long)h * Ml;
        // receive buffer of rank h is ordered like its output window with the local run removed
        const long long r = (h > me) ? k_local : k_local - xplan[XP_RANK_LOC + h];
        unsigned char* out = reinterpret_cast<unsigned char*>(peer_recv[h]) + (size_t)r * stride;
        const double2* src = reinterpret_cast<const double2*>(pose4 + 4 * a);
        double2* dst = reinterpret_cast<double2*>(out);
        dst[0] = src[0];
        dst[1] = src[1];
        const int2 ax = reinterpret_cast<const int2*>(aux2)[a];
        *reinterpret_cast<int2*>(out + 32) = ax;
        src_slot[j] = slot[a];
        dst_idx[j] = (int)r;
        dst_rank[j] = h;
        nlive[j] = ax.x;
    }
}

__global__ void __launch_bounds__(256)
pack_headers_kernel(const long long* __restrict__ emit_run, long long n, long long particle_offset,
                    const double* __restrict__ pose4, const int* __restrict__ aux2, const int* __restrict__ slot,
                    unsigned char* __restrict__ out, long long stride, int* __restrict__ src_slot,
                    int* __restrict__ dst_idx, int* __restrict__ nlive) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a = emit_run[i] - particle_offset;
    const double2* src = reinterpret_cast<const double2*>(pose4 + 4 * a);
    double2* dst = reinterpret_cast<double2*>(out + (size_t)i * stride);
    dst[0] = src[0];
    dst[1] = src[1];
    const int2 ax = reinterpret_cast<const int2*>(aux2)[a];
    *reinterpret_cast<int2*>(out + (size_t)i * stride + 32) = ax;
    src_slot[i] = slot[a];
    dst_idx[i] = (int)i;
    nlive[i] = ax.x;
}

// number of outputs of each local particle that fall inside this rank's own output window
__global__ void __launch_bounds__(256)
offspring_window_kernel(const long long* __restrict__ out_lo, const int* __restrict__ offspring, long long Ml,
                        long long win_lo, int* __restrict__ offspring_local) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ml) return;
    const long long lo = max(out_lo[i], win_lo), hi = min(out_lo[i] + offspring[i], win_lo + Ml);
    offspring_local[i] = (int)max(0ll, hi - lo);
}

// G4s: sharded form of G4.  The rank's Ml output slots are [incoming from lower ranks (n_lo) |
// offspring of local ancestors (n_loc) | incoming from higher ranks (Ml - n_lo - n_loc)] because
// ancestors are globally ascending.  Incoming particles and local duplicates both take blocks
// freed by local particles with no local offspring.
//
// xplan == NULL: n_lo / n_loc are the host's values and local_run[j] is the ancestor of local output
// n_lo + j (NCCL path).  xplan != NULL: the split is read from the device-resident exchange plan and
// local_run is indexed by the output slot itself (peer path; nothing on this path visits the host).
//
// `part`: kAssignAll -- every output slot in one launch (NCCL path: the arrivals are in `recv` already);
// kAssignLocal -- slots filled from local ancestors only; the arrival slots just mark their entry of the
// pool-to-pool copy list as "not a copy", so the local duplicates can be copied while the other ranks' pushes
// are still in flight; kAssignArrivals -- thread i handles arrival i of the receive buffer, after the flag
// barrier (`ps`) that makes every rank's pushes visible.
enum { kAssignAll = 0, kAssignLocal = 1, kAssignArrivals = 2 };

__global__ void __launch_bounds__(256)
assign_sharded_kernel(int part, PeerSync ps, const long long* __restrict__ local_run, long long Ml,
                      long long particle_offset, long long n_lo,
                      long long n_loc, const long long* __restrict__ xplan, const double* __restrict__ pose_in,
                      double* __restrict__ pose_out,
                      const int* __restrict__ aux_in, int* __restrict__ aux_out, const int* __restrict__ slot_in,
                      int* __restrict__ slot_out, const unsigned char* recv, long long stride,
                      const int* __restrict__ dead_excl, const int* __restrict__ free_list,
                      const long long* __restrict__ total_dead, int* __restrict__ copy_src, int* __restrict__ copy_dst,
                      int* __restrict__ copy_nlive, int* __restrict__ unpack_src, int* __restrict__ unpack_dst,
                      int* __restrict__ unpack_nlive) {
    if (ps.flags != nullptr) {
        // sharded filter, peer exchange: the arrivals were pushed into this rank's receive buffer by the other ranks;
        // CTA 0 posts this rank's flag (its own pushes: earlier kernels of this stream, fenced), every CTA waits for
        // all ranks' flags before it reads the buffer
        const unsigned long long t_entry = global_timer_ns();
        peer_sync_thread(ps, (int)threadIdx.x, blockIdx.x == 0 && !ps.posted);
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) peer_sync_account(ps, t_entry);
    }
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ml) return;
    long long run_shift = n_lo;
    if (xplan != nullptr) {
        if (xplan[XP_OVERFLOW] != 0) return;  // exchange capacity exceeded: flagged, nothing is touched
        n_lo = xplan[XP_N_LO];
        n_loc = xplan[XP_N_LOC];
        run_shift = 0;
    }
    const long long n_in = Ml - n_loc;
    const long long n_dups = *total_dead - n_in;  // local outputs that are not the first of their ancestor
    if (part == kAssignArrivals) {
        if (k >= n_in) return;
        if (k >= n_lo) k += n_loc;  // arrival index -> output slot
    }
    double2* dst = reinterpret_cast<double2*>(pose_out + 4 * k);
    if (k < n_lo || k >= n_lo + n_loc) {
        const long long r = (k < n_lo) ? k : k - n_loc;  // index in the receive buffer (source-rank order)
        const long long nidx = (k < n_lo) ? k : n_lo + n_dups + (k - n_lo - n_loc);
        if (part == kAssignLocal) {
            copy_src[nidx] = -1;  // not a pool-to-pool copy
            copy_nlive[nidx] = 0;
            return;
        }
        const double2* src = reinterpret_cast<const double2*>(recv + (size_t)r * stride);
        dst[0] = src[0];
        dst[1] = src[1];
        const int2 ax = *reinterpret_cast<const int2*>(recv + (size_t)r * stride + 32);
        reinterpret_cast<int2*>(aux_out)[k] = ax;
        const int d = free_list[nidx];
        slot_out[k] = d;
        unpack_src[r] = (int)r;
        unpack_dst[r] = d;
        unpack_nlive[r] = ax.x;
        if (part == kAssignAll) {
            copy_src[nidx] = -1;  // not a pool-to-pool copy
            copy_dst[nidx] = d;
            copy_nlive[nidx] = 0;
        }
        return;
    }
    if (part == kAssignArrivals) return;
    const long long a = local_run[k - run_shift] - particle_offset;  // local ancestor
    const double2* src = reinterpret_cast<const double2*>(pose_in + 4 * a);
    dst[0] = src[0];
    dst[1] = src[1];
    const int2 ax = reinterpret_cast<const int2*>(aux_in)[a];
    reinterpret_cast<int2*>(aux_out)[k] = ax;
    const bool first = (k == n_lo) || (local_run[k - run_shift - 1] != local_run[k - run_shift]);
    if (first) {
        slot_out[k] = slot_in[a];
    } else {
        const long long alive_before = a - dead_excl[a];
        const long long nidx = n_lo + (k - n_lo) - alive_before - 1;
        const int d = free_list[nidx];
        slot_out[k] = d;
        copy_src[nidx] = slot_in[a];
        copy_dst[nidx] = d;
        copy_nlive[nidx] = ax.x;
    }
}

// ---------------------------------------------------------------------------------------------
// G5: block mover.  One warp per CTA, one elected lane drives a 4-deep ring of 4 KiB shared
// memory buffers: cp.async.bulk global->shared (mbarrier complete_tx), then cp.async.bulk
// shared->global (bulk group).  The chunk stream runs across items, so the pipeline never drains
// between landmark blocks.
// ---------------------------------------------------------------------------------------------
constexpr int kCopyBufs = 4;
constexpr int kCopyChunk = 4096;

#ifndef PK_COPY_META_CACHE
#define PK_COPY_META_CACHE 1
#endif
__global__ void __launch_bounds__(32)
copy_blocks_kernel(const unsigned char* __restrict__ src_base, unsigned char* __restrict__ dst_base, long long src_stride,
                   long long dst_stride, int hot_b, int cold_b, int capacity, const int* __restrict__ src_slot, const int* __restrict__ dst_slot,
                   const int* __restrict__ nlive, long long n_max, const long long* __restrict__ n_dev,
                   const unsigned long long* __restrict__ dst_tab, const int* __restrict__ dst_rank,
                   const long long* __restrict__ skip_flag, long long orph_off, int orph_len) {
    __shared__ __align__(128) unsigned char buf[kCopyBufs][kCopyChunk];
    __shared__ uint64_t bar[kCopyBufs];
    // The copy list of this CTA's items (source slot, destination slot / rank, live landmarks), 64 items at a time:
    // the lanes that do not drive the bulk copies fetch it 32 items ahead, so the driving lane never waits for a
    // dependent global load between two chunks (with 4 KB blocks that wait was as long as the copy itself).
    __shared__ int m_src[64], m_dst[64], m_nl[64], m_rank[64];
    const int lane = threadIdx.x;
#if !PK_COPY_META_CACHE
    if (lane != 0) return;
#endif
    if (skip_flag != nullptr && *skip_flag != 0) return;  // exchange overflow: the frame's copies are void
    long long n = n_dev ? *n_dev : n_max;
    if (n > n_max) n = n_max;
    if (lane == 0) {
        for (int b = 0; b < kCopyBufs; ++b) mbar_init(&bar[b], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const long long stride = gridDim.x;
    const long long first = blockIdx.x;
    const long long n_mine = first < n ? (n - first + stride - 1) / stride : 0;  // items of this CTA: first + k * stride

#if PK_COPY_META_CACHE
    long long cached = 0;  // items [0, cached) of this CTA have been fetched (ring of 64: the last two batches are valid)
    auto ensure = [&](long long k) {  // warp-uniform
        while (k >= cached && cached < n_mine) {
            const long long kk = cached + lane;
            if (kk < n_mine) {
                const long long item = first + kk * stride;
                const int e = (int)(kk & 63);
                m_src[e] = src_slot[item];
                m_dst[e] = dst_slot[item];
                m_nl[e] = nlive ? min(nlive[item], capacity) : capacity;
                m_rank[e] = dst_tab ? dst_rank[item] : 0;
            }
            cached += 32;
            __syncwarp();
        }
    };
    auto src_of = [&](long long k) { return m_src[k & 63]; };
    auto dst_of = [&](long long k) { return m_dst[k & 63]; };
    auto nl_of = [&](long long k) { return m_nl[k & 63]; };
    auto rank_of = [&](long long k) { return m_rank[k & 63]; };
#else
    auto ensure = [&](long long) {};
    auto src_of = [&](long long k) { return src_slot[first + k * stride]; };
    auto dst_of = [&](long long k) { return dst_slot[first + k * stride]; };
    auto nl_of = [&](long long k) { return nlive ? min(nlive[first + k * stride], capacity) : capacity; };
    auto rank_of = [&](long long k) { return dst_rank[first + k * stride]; };
#endif

    // chunk stream: the issue cursor walks this CTA's items (`k` counts them); what the drain side needs -- where a
    // buffer goes and how many bytes -- is noted per buffer when its load is issued
    __shared__ unsigned long long pend_dst[kCopyBufs];
    __shared__ unsigned pend_bytes[kCopyBufs];
    struct Cursor {
        long long k;
        int seg;        // 0 = hot range, 1 = cold range, 2 = orphan region (spawn mode)
        long long off;  // offset inside the segment
    };
    auto seg_len = [&](long long k, int seg) -> long long {
        if (src_of(k) < 0) return 0;  // entry not served by this launch (e.g. filled from a receive buffer)
        const int nl = nl_of(k);
        // hot keys are 4 B each: round the range up to the 16 B granularity of a bulk copy
        if (seg == 2) return (long long)orph_len;
        return seg == 0 ? (((long long)nl * hot_b + 15) & ~15ll) : (long long)nl * cold_b;
    };
    auto seg_base = [&](int seg) -> long long {
        return seg == 0 ? 0 : (seg == 1 ? (long long)hot_region_bytes(capacity) : orph_off);
    };
    auto advance = [&](Cursor& c) {
        // move to the next chunk, skipping empty segments
        c.off += kCopyChunk;
        while (c.k < n_mine) {
            ensure(c.k);
            if (c.off < seg_len(c.k, c.seg)) break;
            c.off = 0;
            if (++c.seg > 2) {
                c.seg = 0;
                ++c.k;
            }
        }
    };
    Cursor ci{0, 0, -(long long)kCopyChunk};
    advance(ci);
    long long issued = 0, drained = 0;
    while (ci.k < n_mine || drained < issued) {
        // keep up to kCopyBufs-1 loads in flight
        while (ci.k < n_mine && issued < drained + (kCopyBufs - 1)) {
            const int b = (int)(issued % kCopyBufs);
            const long long len = seg_len(ci.k, ci.seg);
            const unsigned bytes = (unsigned)min((long long)kCopyChunk, len - ci.off);
            const long long in_block = seg_base(ci.seg) + ci.off;
            const unsigned char* s = src_base + (size_t)src_of(ci.k) * src_stride + in_block;
            // dst_tab != NULL: the item goes to the receive buffer of rank dst_rank[item] (peer memory) and
            // dst_base only carries the byte offset inside a record (the header size)
            unsigned char* db = dst_tab ? reinterpret_cast<unsigned char*>(dst_tab[rank_of(ci.k)]) +
                                              reinterpret_cast<size_t>(dst_base)
                                        : dst_base;
            unsigned char* d = db + (size_t)dst_of(ci.k) * dst_stride + in_block;
            if (lane == 0) {
                pend_dst[b] = reinterpret_cast<unsigned long long>(d);
                pend_bytes[b] = bytes;
                if (issued >= kCopyBufs) tma_store_wait_read0();  // the store that last read this buffer
                mbar_arrive_expect_tx(&bar[b], bytes);
                tma_load_1d(buf[b], s, bytes, &bar[b]);
            }
            ++issued;
            advance(ci);
        }
        if (drained < issued) {
            const int b = (int)(drained % kCopyBufs);
            if (lane == 0) {
                mbar_wait(&bar[b], (unsigned)((drained / kCopyBufs) & 1));
                tma_store_1d(reinterpret_cast<void*>(pend_dst[b]), buf[b], pend_bytes[b]);
                tma_store_commit();
            }
            ++drained;
        }
    }
    if (lane == 0) tma_store_wait0();
}

// ---------------------------------------------------------------------------------------------
// Log-domain weight normaliser (PK_MODEL_LOG_WEIGHTS): max, then exp(lw - max) with sum and sum of squares.
// Warp-shuffle reductions inside a CTA, one partial per CTA, a fixed-order second pass: deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
logw_max_partial_kernel(const double* __restrict__ pose4, long long M, double* __restrict__ ws) {
    __shared__ double sh[8];
    double m = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
        const double v = pose4[4 * i + 3];
        m = (v > m) ? v : m;   // NaN never wins
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(kFullMask, m, o);
        m = (v > m) ? v : m;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = (sh[w] > m) ? sh[w] : m;
        ws[blockIdx.x] = m;
    }
}

__global__ void logw_max_final_kernel(const double* __restrict__ ws, int nblocks, double* __restrict__ max_out) {
    double m = -INFINITY;
    for (int b = threadIdx.x; b < nblocks; b += 32) m = (ws[b] > m) ? ws[b] : m;
    for (int o = 16; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(kFullMask, m, o);
        m = (v > m) ? v : m;
    }
    if (threadIdx.x == 0) max_out[0] = m;
}

__global__ void __launch_bounds__(256)
logw_normalise_kernel(double* __restrict__ pose4, long long M, const double* __restrict__ max_in, double* __restrict__ ws) {
    __shared__ double sh[2][8];
    const double mx = max_in[0];
    double s = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
        // all weights -inf (every factor was zero): keep them equal, as a uniform resample
        const double w = (mx == -INFINITY) ? 1.0 : pk_exp(pose4[4 * i + 3] - mx);
        pose4[4 * i + 3] = w;
        s += w;
        s2 += w * w;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(kFullMask, s, o);
        s2 += __shfl_xor_sync(kFullMask, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = s;
        sh[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += sh[threadIdx.x][w];
        ws[threadIdx.x * 1024 + blockIdx.x] = a;
    }
}

__global__ void logw_sums_final_kernel(const double* __restrict__ ws, int nblocks, const double* __restrict__ max_in,
                                       double* __restrict__ out3) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;  // two warps: sum, sum of squares
    double a = 0.0;
    for (int b = lane; b < nblocks; b += 32) a += ws[q * 1024 + b];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(kFullMask, a, o);
    if (lane == 0) out3[q] = a;
    if (threadIdx.x == 0) out3[2] = max_in[0];
}

// ---------------------------------------------------------------------------------------------
// K6 summary / best particle
// ---------------------------------------------------------------------------------------------
constexpr int kRedBlocks = 1024;

__global__ void __launch_bounds__(256)
summary_partial_kernel(const double* __restrict__ pose4, long long M, double* __restrict__ ws) {
    __shared__ double sh[4][8];
    double sx = 0.0, sy = 0.0, ss = 0.0, sc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
        const double2 xy = *reinterpret_cast<const double2*>(pose4 + 4 * i);
        const double th = pose4[4 * i + 2];
        double s, c;
        sincos(th, &s, &c);
        sx += xy.x;  // :267-271
        sy += xy.y;
        ss += s;
        sc += c;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(kFullMask, sx, o);
        sy += __shfl_xor_sync(kFullMask, sy, o);
        ss += __shfl_xor_sync(kFullMask, ss, o);
        sc += __shfl_xor_sync(kFullMask, sc, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sh[0][warp] = sx;
        sh[1][warp] = sy;
        sh[2][warp] = ss;
        sh[3][warp] = sc;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += sh[threadIdx.x][w];
        ws[threadIdx.x * kRedBlocks + blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(128)
summary_final_kernel(const double* __restrict__ ws, int nblocks, long long M, double* __restrict__ out5) {
    // four warps, one per quantity; fixed-order reduction of the per-block partials
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double a = 0.0;
    for (int b = lane; b < nblocks; b += 32) a += ws[q * kRedBlocks + b];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(kFullMask, a, o);
    if (lane == 0) out5[q] = a;
    if (threadIdx.x == 0) out5[4] = (double)M;
}

__global__ void __launch_bounds__(256)
best_partial_kernel(const double* __restrict__ pose4, long long M, double* __restrict__ ws) {
    __shared__ double shv[8];
    __shared__ long long shi[8];
    double bv = -1.0;
    long long bi = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
        const double w = pose4[4 * i + 3];
        if (w > bv) {  // strict: first maximum wins within a thread (indices ascend)
            bv = w;
            bi = i;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(kFullMask, bv, o);
        const long long oi = __shfl_xor_sync(kFullMask, bi, o);
        if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) {
            bv = ov;
            bi = oi;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        shv[warp] = bv;
        shi[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (shv[w] > bv || (shv[w] == bv && shi[w] >= 0 && (bi < 0 || shi[w] < bi))) {
                bv = shv[w];
                bi = shi[w];
            }
        ws[blockIdx.x] = bv;
        ws[kRedBlocks + blockIdx.x] = (double)bi;
    }
}

__global__ void best_final_kernel(const double* __restrict__ ws, int nblocks, double* __restrict__ best2) {
    if (threadIdx.x != 0) return;
    double bv = -1.0, bi = -1.0;
    for (int b = 0; b < nblocks; ++b) {
        const double v = ws[b], i = ws[kRedBlocks + b];
        if (v > bv || (v == bv && i >= 0.0 && (bi < 0.0 || i < bi))) {
            bv = v;
            bi = i;
        }
    }
    best2[0] = bv;
    best2[1] = bi;
}

static long long num_blocks(long long M) { return (M + PK_SCAN_BLOCK - 1) / PK_SCAN_BLOCK; }

struct GatherWs {
    int *dead_excl, *block_dead, *block_off, *free_list, *copy_src, *copy_dst, *copy_nlive;
    int *offspring_local, *unpack_src, *unpack_dst, *unpack_nlive;
    size_t bytes;
};
static GatherWs carve(void* ws, long long M) {
    GatherWs g;
    const long long nb = num_blocks(M);
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    unsigned char* base = (unsigned char*)ws;
    auto take = [&](size_t n) {
        int* p = (int*)(base + off);
        off += align(n * sizeof(int));
        return p;
    };
    g.dead_excl = take((size_t)M);
    g.block_dead = take((size_t)nb);
    g.block_off = take((size_t)nb + 1);
    g.free_list = take((size_t)M);
    g.copy_src = take((size_t)M);
    g.copy_dst = take((size_t)M);
    g.copy_nlive = take((size_t)M);
    g.offspring_local = take((size_t)M);
    g.unpack_src = take((size_t)M);
    g.unpack_dst = take((size_t)M);
    g.unpack_nlive = take((size_t)M);
    g.bytes = off;
    return g;
}

}  // namespace pk

using namespace pk;

extern "C" {

long long pk_num_scan_blocks(long long M) { return num_blocks(M); }

int pk_weight_scan(const double* pose4, long long M, double* cumsum, double* block_sums, void* stream) {
    PK_CHECK_ARG(pose4 && cumsum && block_sums, "null pointer");
    PK_CHECK_ARG(M > 0, "M <= 0");
    const long long nb = num_blocks(M);
    const long long grid = (nb + kScanWarps - 1) / kScanWarps;
    weight_scan_kernel<<<(unsigned)grid, kScanWarps * 32, 0, (cudaStream_t)stream>>>(pose4, M, cumsum, block_sums, nb,
                                                                                     nullptr, 0, 0);
    PK_LAUNCH_CHECK("weight_scan_kernel");
    return PK_OK;
}

int pk_resample_thresholds(const double* all_block_sums, long long nb_total, long long M_total, double u01, double* plan,
                           double* block_prefix, long long* block_count, void* stream) {
    PK_CHECK_ARG(all_block_sums && plan && block_prefix && block_count, "null pointer");
    PK_CHECK_ARG(nb_total > 0 && M_total > 0, "sizes");
    PK_CHECK_ARG(nb_total <= (long long)kMaxGroupBlocks * kMaxScanGroups, "more than 2^25 particles in one filter");
    // the fold tree depends on the TOTAL block count only, never on how the blocks are spread over ranks
    int group = kMinGroupBlocks;
    while ((long long)group * kMaxScanGroups < nb_total) group *= 2;
    thresholds_kernel<<<kThrCluster, 1024, 0, (cudaStream_t)stream>>>(all_block_sums, nb_total, M_total, u01, plan, block_prefix,
                                                           block_count, group, PeerSync{nullptr, 0, 1, 0, 0, nullptr, 0, false}, 0, 0,
                                                           nullptr);
    PK_LAUNCH_CHECK("thresholds_kernel");
    return PK_OK;
}

int pk_resample_thresholds_peer(const double* all_block_sums, long long nb_total, long long M_total, double u01,
                                double* plan, double* block_prefix, long long* block_count,
                                const unsigned long long* peer_flags_tab, int rank, int n_ranks, unsigned long long epoch,
                                double timeout_s, long long Ml, long long capacity, long long* xplan,
                                unsigned long long* status, void* stream) {
    PK_CHECK_ARG(all_block_sums && plan && block_prefix && block_count && peer_flags_tab && xplan && status, "null pointer");
    PK_CHECK_ARG(nb_total > 0 && M_total > 0 && Ml > 0 && capacity >= 0, "sizes");
    PK_CHECK_ARG(nb_total <= (long long)kMaxGroupBlocks * kMaxScanGroups, "more than 2^25 particles in one filter");
    PK_CHECK_ARG(n_ranks >= 1 && n_ranks <= PK_MAX_RANKS && rank >= 0 && rank < n_ranks && nb_total % n_ranks == 0,
                 "rank / n_ranks");
    PK_CHECK_ARG(epoch > 0 && timeout_s > 0.0, "barrier arguments");
    int group = kMinGroupBlocks;
    while ((long long)group * kMaxScanGroups < nb_total) group *= 2;
    thresholds_kernel<<<kThrCluster, 1024, 0, (cudaStream_t)stream>>>(
        all_block_sums, nb_total, M_total, u01, plan, block_prefix, block_count, group,
        PeerSync{reinterpret_cast<unsigned long long* const*>(peer_flags_tab), rank, n_ranks, epoch,
                 (unsigned long long)(timeout_s * 1e9), status, 0, false},
        Ml, capacity, xplan);
    PK_LAUNCH_CHECK("thresholds_kernel");
    return PK_OK;
}

int pk_resample_ancestors(const double* cumsum, long long M_local, long long particle_offset, long long block_offset,
                          const double* plan, const double* block_prefix, const long long* block_count,
                          long long M_total, long long out_offset, long long n_out, long long* out_lo, int* offspring,
                          long long* ancestors, long long* big_runs, void* stream) {
    PK_CHECK_ARG(cumsum && plan && block_prefix && block_count && out_lo && offspring && ancestors && big_runs,
                 "null pointer");
    PK_CHECK_ARG(M_local > 0 && M_total >= M_local, "sizes");
    PK_CHECK_ARG(particle_offset % PK_SCAN_BLOCK == 0, "particle_offset must be a multiple of PK_SCAN_BLOCK");
    cudaStream_t st = (cudaStream_t)stream;
    PK_CUDA(cudaMemsetAsync(big_runs, 0, 4 * sizeof(long long), st));
    const int threads = 256;
    ancestors_kernel<<<(unsigned)((M_local + threads - 1) / threads), threads, 0, st>>>(
        cumsum, M_local, particle_offset, block_offset, plan, block_prefix, block_count, M_total, out_offset, n_out,
        out_lo, offspring, ancestors, big_runs);
    PK_LAUNCH_CHECK("ancestors_kernel");
    fill_runs_kernel<<<num_sms() * 2, 256, 0, st>>>(big_runs, ancestors);
    PK_LAUNCH_CHECK("fill_runs_kernel");
    return PK_OK;
}

long long pk_gather_workspace_bytes(long long M) {
    if (M <= 0) return 0;
    return (long long)carve(nullptr, M).bytes;
}

static int copy_blocks_launch(const void* src, void* dst, int capacity, int dtype, const int* src_slot,
                              const int* dst_slot, const int* nlive, long long n_max, const long long* n_dev,
                              cudaStream_t st, long long src_stride = 0, long long dst_stride = 0,
                              const unsigned long long* dst_tab = nullptr, const int* dst_rank = nullptr,
                              const long long* skip_flag = nullptr) {
    if (src_stride == 0) src_stride = (long long)block_bytes(capacity, dtype);
    if (dst_stride == 0) dst_stride = (long long)block_bytes(capacity, dtype);
    long long grid = (long long)num_sms() * 12;
    if (grid > n_max) grid = n_max;
    if (grid < 1) return PK_OK;
    copy_blocks_kernel<<<(unsigned)grid, 32, 0, st>>>((const unsigned char*)src, (unsigned char*)dst, src_stride,
                                                     dst_stride, (int)hot_bytes(dtype),
                                                     (int)cold_bytes(dtype), capacity, src_slot, dst_slot, nlive, n_max,
                                                     n_dev, dst_tab, dst_rank, skip_flag,
                                                     (long long)orphan_offset(capacity, dtype),
                                                     (int)orphan_region_bytes(dtype));
    PK_LAUNCH_CHECK("copy_blocks_kernel");
    return PK_OK;
}

int pk_resample_gather(const long long* ancestors, const int* offspring, long long M, const double* pose4_in,
                       double* pose4_out, const int* aux2_in, int* aux2_out, const int* slot_in, int* slot_out,
                       void* pool, int capacity, int dtype, void* workspace, long long* n_copied_out, void* stream) {
    PK_CHECK_ARG(ancestors && offspring && pose4_in && pose4_out && aux2_in && aux2_out && slot_in && slot_out && pool &&
                     workspace && n_copied_out,
                 "null pointer");
    PK_CHECK_ARG(M > 0 && M < (1ll << 31), "M");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    PK_CHECK_ARG(pose4_in != pose4_out && slot_in != slot_out && aux2_in != aux2_out, "gather is out of place");
    cudaStream_t st = (cudaStream_t)stream;
    GatherWs g = carve(workspace, M);
    const long long nb = num_blocks(M);
    dead_scan_kernel<<<(unsigned)((nb + kScanWarps - 1) / kScanWarps), kScanWarps * 32, 0, st>>>(offspring, M, g.dead_excl,
                                                                                                 g.block_dead, nb);
    PK_LAUNCH_CHECK("dead_scan_kernel");
    block_offsets_kernel<<<1, 1024, 0, st>>>(g.block_dead, nb, g.block_off, n_copied_out);
    PK_LAUNCH_CHECK("block_offsets_kernel");
    const int threads = 256;
    const unsigned grid = (unsigned)((M + threads - 1) / threads);
    free_list_kernel<<<grid, threads, 0, st>>>(offspring, slot_in, M, g.dead_excl, g.block_off, g.free_list);
    PK_LAUNCH_CHECK("free_list_kernel");
    assign_kernel<<<grid, threads, 0, st>>>(ancestors, M, pose4_in, pose4_out, aux2_in, aux2_out, slot_in, slot_out,
                                            g.dead_excl, g.free_list, g.copy_src, g.copy_dst, g.copy_nlive);
    PK_LAUNCH_CHECK("assign_kernel");
    if (capacity > 0)
        return copy_blocks_launch(pool, pool, capacity, dtype, g.copy_src, g.copy_dst, g.copy_nlive, M, n_copied_out, st);
    return PK_OK;
}

int pk_resample_plan(const double* cumsum, long long M_local, long long particle_offset, long long block_offset,
                     const double* plan, const double* block_prefix, const long long* block_count, long long M_total,
                     long long out_offset, long long n_out, long long* out_lo, int* offspring, long long* ancestors,
                     void* gather_workspace, void* stream) {
    PK_CHECK_ARG(cumsum && plan && block_prefix && block_count && out_lo && offspring && ancestors && gather_workspace,
                 "null pointer");
    PK_CHECK_ARG(M_local > 0 && M_local < (1ll << 31) && M_total >= M_local, "sizes");
    PK_CHECK_ARG(particle_offset % PK_SCAN_BLOCK == 0, "particle_offset must be a multiple of PK_SCAN_BLOCK");
    GatherWs g = carve(gather_workspace, M_local);
    const long long nb = num_blocks(M_local);
    resample_plan_kernel<<<(unsigned)nb, PK_SCAN_BLOCK, 0, (cudaStream_t)stream>>>(
        cumsum, M_local, particle_offset, block_offset, plan, block_prefix, block_count, M_total, out_offset, n_out, out_lo,
        offspring, ancestors, g.offspring_local, g.dead_excl, g.block_dead);
    PK_LAUNCH_CHECK("resample_plan_kernel");
    return PK_OK;
}

int pk_resample_gather_planned(const long long* ancestors, long long M, const double* pose4_in, double* pose4_out,
                               const int* aux2_in, int* aux2_out, const int* slot_in, int* slot_out, void* pool,
                               int capacity, int dtype, void* workspace, long long* n_copied_out, void* stream) {
    PK_CHECK_ARG(ancestors && pose4_in && pose4_out && aux2_in && aux2_out && slot_in && slot_out && pool && workspace &&
                     n_copied_out,
                 "null pointer");
    PK_CHECK_ARG(M > 0 && M < (1ll << 31), "M");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    PK_CHECK_ARG(pose4_in != pose4_out && slot_in != slot_out && aux2_in != aux2_out, "gather is out of place");
    cudaStream_t st = (cudaStream_t)stream;
    GatherWs g = carve(workspace, M);
    const long long nb = num_blocks(M);
    free_list_fused_kernel<<<(unsigned)nb, PK_SCAN_BLOCK, 0, st>>>(g.offspring_local, slot_in, M, g.dead_excl, g.block_dead, nb,
                                                                  g.free_list, n_copied_out);
    PK_LAUNCH_CHECK("free_list_fused_kernel");
    const int threads = 256;
    assign_kernel<<<(unsigned)((M + threads - 1) / threads), threads, 0, st>>>(
        ancestors, M, pose4_in, pose4_out, aux2_in, aux2_out, slot_in, slot_out, g.dead_excl, g.free_list, g.copy_src,
        g.copy_dst, g.copy_nlive);
    PK_LAUNCH_CHECK("assign_kernel");
    if (capacity > 0)
        return copy_blocks_launch(pool, pool, capacity, dtype, g.copy_src, g.copy_dst, g.copy_nlive, M, n_copied_out, st);
    return PK_OK;
}

int pk_resample_copy_blocks(void* pool, int capacity, int dtype, long long M, void* workspace,
                            const long long* n_copied, void* stream) {
    PK_CHECK_ARG(pool && workspace && n_copied, "null pointer");
    PK_CHECK_ARG(M > 0 && M < (1ll << 31), "M");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    if (capacity <= 0) return PK_OK;
    GatherWs g = carve(workspace, M);
    return copy_blocks_launch(pool, pool, capacity, dtype, g.copy_src, g.copy_dst, g.copy_nlive, M, n_copied,
                              (cudaStream_t)stream);
}

long long pk_particle_record_bytes(int capacity, int dtype) {
    return (long long)kHeaderBytes + (long long)block_bytes(capacity, dtype);
}

int pk_pack_particles(const long long* emit_run, long long n, long long particle_offset, const double* pose4,
                      const int* aux2, const int* slot, const void* pool, int capacity, int dtype, void* out,
                      int* workspace, void* stream) {
    PK_CHECK_ARG(n >= 0, "n < 0");
    if (n == 0) return PK_OK;
    PK_CHECK_ARG(emit_run && pose4 && aux2 && slot && pool && out && workspace, "null pointer");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const long long stride = pk_particle_record_bytes(capacity, dtype);
    int* src_slot = workspace;
    int* dst_idx = workspace + n;
    int* nlive = workspace + 2 * n;
    pack_headers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(emit_run, n, particle_offset, pose4, aux2, slot,
                                                                    (unsigned char*)out, stride, src_slot, dst_idx, nlive);
    PK_LAUNCH_CHECK("pack_headers_kernel");
    if (capacity > 0)
        return copy_blocks_launch(pool, (unsigned char*)out + kHeaderBytes, capacity, dtype, src_slot, dst_idx, nlive, n,
                                  nullptr, st, 0, stride);
    return PK_OK;
}

static int gather_sharded_impl(bool planned, PeerSync sync, const long long* local_run, const long long* xplan, long long recv_capacity,
                               const long long* out_lo, const int* offspring, long long Ml, long long particle_offset,
                               long long n_lo, long long n_loc, const double* pose4_in, double* pose4_out,
                               const int* aux2_in, int* aux2_out, const int* slot_in, int* slot_out, const void* recv,
                               void* pool, int capacity, int dtype, void* workspace, long long* total_dead_out,
                               cudaStream_t st, cudaEvent_t pushes_done = nullptr) {
    GatherWs g = carve(workspace, Ml);
    const long long nb = num_blocks(Ml);
    const long long stride = pk_particle_record_bytes(capacity, dtype);
    const int threads = 256;
    const unsigned grid = (unsigned)((Ml + threads - 1) / threads);
    if (!planned) {  // pk_resample_plan has not run for this window: windowed offspring and the dead scan, separately
        offspring_window_kernel<<<grid, threads, 0, st>>>(out_lo, offspring, Ml, particle_offset, g.offspring_local);
        PK_LAUNCH_CHECK("offspring_window_kernel");
        dead_scan_kernel<<<(unsigned)((nb + kScanWarps - 1) / kScanWarps), kScanWarps * 32, 0, st>>>(
            g.offspring_local, Ml, g.dead_excl, g.block_dead, nb);
        PK_LAUNCH_CHECK("dead_scan_kernel");
    }
    const PeerSync no_sync{nullptr, 0, 1, 0, 0, nullptr, 0};
    free_list_fused_kernel<<<(unsigned)nb, PK_SCAN_BLOCK, 0, st>>>(g.offspring_local, slot_in, Ml, g.dead_excl, g.block_dead,
                                                                  nb, g.free_list, total_dead_out);
    PK_LAUNCH_CHECK("free_list_fused_kernel");
    auto assign = [&](int part, const PeerSync& ps) {
        assign_sharded_kernel<<<grid, threads, 0, st>>>(part, ps, local_run, Ml, particle_offset, n_lo, n_loc, xplan, pose4_in,
                                                        pose4_out, aux2_in, aux2_out, slot_in, slot_out,
                                                        (const unsigned char*)recv, stride, g.dead_excl, g.free_list,
                                                        total_dead_out, g.copy_src, g.copy_dst, g.copy_nlive, g.unpack_src,
                                                        g.unpack_dst, g.unpack_nlive);
    };
    // peer exchange: everything that does not need the arrivals runs first -- the other ranks' pushes (and this
    // rank's own) are still draining over NVLink while the local duplicates are copied at HBM speed; the flag barrier
    // sits in front of the arrivals' part
    const bool split = sync.flags != nullptr;
    assign(split ? kAssignLocal : kAssignAll, no_sync);
    PK_LAUNCH_CHECK("assign_sharded_kernel");
    const long long* skip = xplan ? xplan + XP_OVERFLOW : nullptr;
    // this rank's pushes were launched on another stream: a particle whose offspring all live on other ranks is dead
    // here, and its block -- which the push still reads -- may be handed to a local duplicate; the flag must not be
    // posted before the pushes are complete either
    if (pushes_done != nullptr) PK_CUDA(cudaStreamWaitEvent(st, pushes_done, 0));
    if (capacity > 0) {
        // local duplicates: pool -> pool (entries that belong to incoming particles carry src = -1 and are skipped)
        int rc = copy_blocks_launch(pool, pool, capacity, dtype, g.copy_src, g.copy_dst, g.copy_nlive, Ml, total_dead_out, st,
                                    0, 0, nullptr, nullptr, skip);
        if (rc != PK_OK) return rc;
    }
    if (split) {
        assign(kAssignArrivals, sync);
        PK_LAUNCH_CHECK("assign_sharded_kernel(arrivals)");
    }
    if (capacity > 0) {
        // arrivals: exchange buffer -> pool
        if (xplan != nullptr) {
            const long long n_max = recv_capacity < Ml ? recv_capacity : Ml;
            if (n_max > 0)
                return copy_blocks_launch((const unsigned char*)recv + kHeaderBytes, pool, capacity, dtype, g.unpack_src,
                                          g.unpack_dst, g.unpack_nlive, n_max, xplan + XP_N_IN, st, stride, 0, nullptr,
                                          nullptr, skip);
            return PK_OK;
        }
        const long long n_in = Ml - n_loc;
        if (n_in > 0)
            return copy_blocks_launch((const unsigned char*)recv + kHeaderBytes, pool, capacity, dtype, g.unpack_src,
                                      g.unpack_dst, g.unpack_nlive, n_in, nullptr, st, stride, 0);
    }
    return PK_OK;
}

int pk_resample_gather_sharded(const long long* local_run, const long long* out_lo, const int* offspring, long long Ml,
                               long long particle_offset, long long n_lo, long long n_loc, const double* pose4_in,
                               double* pose4_out, const int* aux2_in, int* aux2_out, const int* slot_in, int* slot_out,
                               const void* recv, void* pool, int capacity, int dtype, void* workspace,
                               long long* total_dead_out, void* stream) {
    PK_CHECK_ARG(out_lo && offspring && pose4_in && pose4_out && aux2_in && aux2_out && slot_in && slot_out && pool &&
                     workspace && total_dead_out,
                 "null pointer");
    PK_CHECK_ARG(Ml > 0 && Ml < (1ll << 31), "Ml");
    PK_CHECK_ARG(n_lo >= 0 && n_loc >= 0 && n_lo + n_loc <= Ml, "window split");
    PK_CHECK_ARG(n_loc == 0 || local_run != nullptr, "local_run is NULL");
    PK_CHECK_ARG(n_loc == Ml || recv != nullptr, "receive buffer is NULL");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    return gather_sharded_impl(false, PeerSync{nullptr, 0, 1, 0, 0, nullptr, 0, false}, local_run, nullptr, 0, out_lo, offspring, Ml, particle_offset, n_lo, n_loc, pose4_in,
                               pose4_out, aux2_in, aux2_out, slot_in, slot_out, recv, pool, capacity, dtype, workspace,
                               total_dead_out, (cudaStream_t)stream);
}

int pk_copy_blocks(const void* pool_src, void* pool_dst, int capacity, int dtype, const int* src_slot,
                   const int* dst_slot, const int* n_live, long long n_max, const long long* n_dev, void* stream) {
    PK_CHECK_ARG(pool_src && pool_dst && src_slot && dst_slot, "null pointer");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    PK_CHECK_ARG(capacity > 0 && n_max >= 0, "sizes");
    if (n_max == 0) return PK_OK;
    return copy_blocks_launch(pool_src, pool_dst, capacity, dtype, src_slot, dst_slot, n_live, n_max, n_dev,
                              (cudaStream_t)stream);
}

int pk_log_weights_max(const double* pose4, long long M, double* max_out, double* workspace, void* stream) {
    PK_CHECK_ARG(pose4 && max_out && workspace, "null pointer");
    PK_CHECK_ARG(M > 0, "M <= 0");
    long long blocks = (M + 255) / 256;
    if (blocks > kRedBlocks) blocks = kRedBlocks;
    cudaStream_t st = (cudaStream_t)stream;
    logw_max_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(pose4, M, workspace);
    PK_LAUNCH_CHECK("logw_max_partial_kernel");
    logw_max_final_kernel<<<1, 32, 0, st>>>(workspace, (int)blocks, max_out);
    PK_LAUNCH_CHECK("logw_max_final_kernel");
    return PK_OK;
}

int pk_log_weights_normalise(double* pose4, long long M, const double* max_in, double* out3, double* workspace,
                             void* stream) {
    PK_CHECK_ARG(pose4 && max_in && out3 && workspace, "null pointer");
    PK_CHECK_ARG(M > 0, "M <= 0");
    long long blocks = (M + 255) / 256;
    if (blocks > kRedBlocks) blocks = kRedBlocks;
    cudaStream_t st = (cudaStream_t)stream;
    logw_normalise_kernel<<<(unsigned)blocks, 256, 0, st>>>(pose4, M, max_in, workspace);
    PK_LAUNCH_CHECK("logw_normalise_kernel");
    logw_sums_final_kernel<<<1, 64, 0, st>>>(workspace, (int)blocks, max_in, out3);
    PK_LAUNCH_CHECK("logw_sums_final_kernel");
    return PK_OK;
}

int pk_summary_partial(const double* pose4, long long M, double* out5, double* workspace, void* stream) {
    PK_CHECK_ARG(pose4 && out5 && workspace, "null pointer");
    PK_CHECK_ARG(M > 0, "M <= 0");
    long long blocks = (M + 255) / 256;
    if (blocks > kRedBlocks) blocks = kRedBlocks;
    cudaStream_t st = (cudaStream_t)stream;
    summary_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(pose4, M, workspace);
    PK_LAUNCH_CHECK("summary_partial_kernel");
    summary_final_kernel<<<1, 128, 0, st>>>(workspace, (int)blocks, M, out5);
    PK_LAUNCH_CHECK("summary_final_kernel");
    return PK_OK;
}

int pk_best_particle(const double* pose4, long long M, double* best2, double* workspace, void* stream) {
    PK_CHECK_ARG(pose4 && best2 && workspace, "null pointer");
    PK_CHECK_ARG(M > 0, "M <= 0");
    long long blocks = (M + 255) / 256;
    if (blocks > kRedBlocks) blocks = kRedBlocks;
    cudaStream_t st = (cudaStream_t)stream;
    best_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(pose4, M, workspace);
    PK_LAUNCH_CHECK("best_partial_kernel");
    best_final_kernel<<<1, 32, 0, st>>>(workspace, (int)blocks, best2);
    PK_LAUNCH_CHECK("best_final_kernel");
    return PK_OK;
}

/* ---- peer path: the same resampling with every count resident on the device and the exchange done by
 *      this library's own kernels over NVLink peer memory (no NCCL call, no host round trip) -------- */
int pk_weight_scan_publish(const double* pose4, long long M, double* cumsum, double* block_sums,
                           const unsigned long long* peer_sums_tab, int rank, int n_ranks, void* stream) {
    PK_CHECK_ARG(pose4 && cumsum && block_sums && peer_sums_tab, "null pointer");
    PK_CHECK_ARG(M > 0, "M <= 0");
    PK_CHECK_ARG(n_ranks >= 1 && n_ranks <= PK_MAX_RANKS && rank >= 0 && rank < n_ranks, "rank / n_ranks");
    const long long nb = num_blocks(M);
    const long long grid = (nb + kScanWarps - 1) / kScanWarps;
    weight_scan_kernel<<<(unsigned)grid, kScanWarps * 32, 0, (cudaStream_t)stream>>>(
        pose4, M, cumsum, block_sums, nb, reinterpret_cast<double* const*>(peer_sums_tab), rank, n_ranks);
    PK_LAUNCH_CHECK("weight_scan_kernel");
    return PK_OK;
}

int pk_peer_barrier(const unsigned long long* peer_flags_tab, int rank, int n_ranks, unsigned long long epoch,
                    double timeout_s, unsigned long long* status, void* stream) {
    PK_CHECK_ARG(peer_flags_tab && status, "null pointer");
    PK_CHECK_ARG(n_ranks >= 1 && n_ranks <= PK_MAX_RANKS && rank >= 0 && rank < n_ranks, "rank / n_ranks");
    PK_CHECK_ARG(epoch > 0, "epoch must start at 1");
    PK_CHECK_ARG(timeout_s > 0.0, "timeout");
    peer_barrier_kernel<<<1, PK_MAX_RANKS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<unsigned long long* const*>(peer_flags_tab), rank, n_ranks, epoch,
        (unsigned long long)(timeout_s * 1e9), status);
    PK_LAUNCH_CHECK("peer_barrier_kernel");
    return PK_OK;
}

int pk_peer_post(const unsigned long long* peer_flags_tab, int rank, int n_ranks, unsigned long long epoch, void* stream) {
    PK_CHECK_ARG(peer_flags_tab != nullptr, "null pointer");
    PK_CHECK_ARG(n_ranks >= 1 && n_ranks <= PK_MAX_RANKS && rank >= 0 && rank < n_ranks, "rank / n_ranks");
    PK_CHECK_ARG(epoch > 0, "epoch must start at 1");
    peer_post_kernel<<<1, PK_MAX_RANKS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<unsigned long long* const*>(peer_flags_tab), rank, n_ranks, epoch);
    PK_LAUNCH_CHECK("peer_post_kernel");
    return PK_OK;
}

int pk_exchange_plan(const long long* block_count, long long nb_per_rank, int n_ranks, int rank, long long Ml,
                     long long capacity, long long* xplan, unsigned long long* status, void* stream) {
    PK_CHECK_ARG(block_count && xplan && status, "null pointer");
    PK_CHECK_ARG(n_ranks >= 1 && n_ranks <= PK_MAX_RANKS && rank >= 0 && rank < n_ranks, "rank / n_ranks");
    PK_CHECK_ARG(nb_per_rank > 0 && Ml > 0 && capacity >= 0, "sizes");
    exchange_plan_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(block_count, nb_per_rank, n_ranks, rank, Ml, capacity, xplan,
                                                             status);
    PK_LAUNCH_CHECK("exchange_plan_kernel");
    return PK_OK;
}

int pk_exchange_plan_host(const long long* emitted_before_host, int n_ranks, int rank, long long Ml, long long capacity,
                          long long* xplan_host) {
    PK_CHECK_ARG(emitted_before_host && xplan_host, "null pointer");
    PK_CHECK_ARG(n_ranks >= 1 && n_ranks <= PK_MAX_RANKS && rank >= 0 && rank < n_ranks, "rank / n_ranks");
    PK_CHECK_ARG(Ml > 0 && capacity >= 0, "sizes");
    make_exchange_plan(emitted_before_host, n_ranks, Ml, rank, capacity, xplan_host);
    return PK_OK;
}

int pk_push_particles(const long long* xplan, const long long* out_lo, long long Ml, int rank, const double* pose4,
                      const int* aux2, const int* slot, const void* pool, int capacity, int dtype,
                      const unsigned long long* peer_recv_tab, long long send_capacity, int* workspace, void* stream) {
    PK_CHECK_ARG(xplan && out_lo && pose4 && aux2 && slot && pool && peer_recv_tab && workspace, "null pointer");
    PK_CHECK_ARG(Ml > 0 && send_capacity >= 0 && send_capacity < (1ll << 31), "sizes");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    if (send_capacity == 0) return PK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long stride = pk_particle_record_bytes(capacity, dtype);
    int* src_slot = workspace;
    int* dst_idx = workspace + send_capacity;
    int* dst_rank = workspace + 2 * send_capacity;
    int* nlive = workspace + 3 * send_capacity;
    long long grid = (send_capacity + 255) / 256;
    if (grid > 4ll * num_sms()) grid = 4ll * num_sms();
    push_headers_kernel<<<(unsigned)grid, 256, 0, st>>>(xplan, out_lo, Ml, rank, pose4, aux2, slot, peer_recv_tab, stride,
                                                        send_capacity, src_slot, dst_idx, dst_rank, nlive);
    PK_LAUNCH_CHECK("push_headers_kernel");
    if (capacity > 0)
        return copy_blocks_launch(pool, reinterpret_cast<void*>((size_t)kHeaderBytes), capacity, dtype, src_slot, dst_idx,
                                  nlive, send_capacity, xplan + XP_N_SEND, st, 0, stride, peer_recv_tab, dst_rank);
    return PK_OK;
}

int pk_resample_gather_peer(const long long* xplan, const long long* anc_window, const long long* out_lo,
                            const int* offspring, long long Ml, long long particle_offset, const double* pose4_in,
                            double* pose4_out, const int* aux2_in, int* aux2_out, const int* slot_in, int* slot_out,
                            const void* recv, long long recv_capacity, void* pool, int capacity, int dtype,
                            void* workspace, long long* total_dead_out, const unsigned long long* peer_flags_tab,
                            int rank, int n_ranks, unsigned long long epoch, double timeout_s,
                            unsigned long long* status, void* pushes_done_event, void* stream) {
    PK_CHECK_ARG(xplan && anc_window && out_lo && offspring && pose4_in && pose4_out && aux2_in && aux2_out && slot_in &&
                     slot_out && pool && workspace && total_dead_out && recv,
                 "null pointer");
    PK_CHECK_ARG(Ml > 0 && Ml < (1ll << 31) && recv_capacity >= 0, "sizes");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    PeerSync sync{nullptr, 0, 1, 0, 0, nullptr, 0};
    if (peer_flags_tab != nullptr) {
        PK_CHECK_ARG(n_ranks >= 1 && n_ranks <= PK_MAX_RANKS && rank >= 0 && rank < n_ranks, "rank / n_ranks");
        PK_CHECK_ARG(epoch > 0 && timeout_s > 0.0 && status != nullptr, "barrier arguments");
        sync = PeerSync{reinterpret_cast<unsigned long long* const*>(peer_flags_tab), rank, n_ranks, epoch,
                        (unsigned long long)(timeout_s * 1e9), status, 1, pushes_done_event != nullptr};
    }
    return gather_sharded_impl(true, sync, anc_window, xplan, recv_capacity, out_lo, offspring, Ml, particle_offset, 0, 0, pose4_in,
                               pose4_out, aux2_in, aux2_out, slot_in, slot_out, recv, pool, capacity, dtype, workspace,
                               total_dead_out, (cudaStream_t)stream, (cudaEvent_t)pushes_done_event);
}

/* peer memory: plain cudaMalloc allocations shared between the ranks of one node by CUDA IPC */
int pk_peer_alloc(long long bytes, void** ptr_out) {
    PK_CHECK_ARG(ptr_out && bytes > 0, "arguments");
    void* p = nullptr;
    PK_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "cudaMemset(peer buffer)");
    }
    *ptr_out = p;
    return PK_OK;
}

int pk_peer_free(void* ptr) {
    if (ptr) PK_CUDA(cudaFree(ptr));
    return PK_OK;
}

int pk_peer_export(void* ptr, unsigned char* handle_out_host) {
    PK_CHECK_ARG(ptr && handle_out_host, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == PK_PEER_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    PK_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out_host, &h, sizeof(h));
    return PK_OK;
}

int pk_peer_open(const unsigned char* handle_host, void** ptr_out) {
    PK_CHECK_ARG(handle_host && ptr_out, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    void* p = nullptr;
    PK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return PK_OK;
}

int pk_peer_close(void* ptr) {
    if (ptr) PK_CUDA(cudaIpcCloseMemHandle(ptr));
    return PK_OK;
}

}  // extern "C"
