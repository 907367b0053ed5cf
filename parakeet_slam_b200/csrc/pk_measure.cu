// K2 -- fused association + EKF landmark update + importance weight.
//
// Replaces the per-particle body of FastSLAM.cam_cb (reference prkt_core_v2.py:84-124):
//   match_features_to_scan / match_one (:317-381), probability_of_match (:383-455),
//   prob_position_match (:457-494), closest_point (:496-522, utils.py:37-81), prob_color_match
//   (:524-544, scipy multivariate_normal.pdf), generate_measurement (:859-877),
//   measurement_jacobian (:748-802), measurement_covariance (:804-819), matrix.inverse
//   (matrix.py:11), kalman_gain (:821-833), Feature.update_mean/update_covar (:897-930),
//   importance_factor (:835-849), no_match_weight (:851-857), promotion (:109-118) and the
//   next_id bump of add_orphaned_reading (:740-746).
//
// Shape of the kernel (persistent, one warp owns a *group* of consecutive particles):
//   * the "hot" part of each particle's map (colour mean + meta, 16 B / landmark in f32) is
//     streamed global -> shared by 1-D TMA bulk copies (cp.async.bulk + mbarrier) through a
//     per-warp ring of tiles, several tiles ahead of the consumer;
//   * lanes stride the landmarks of a tile and apply the colour gate (:441) to all K blobs -- the
//     cheapest and most selective of the reference's gates, and the result of
//     probability_of_match is 0 whenever it fails, whatever the evaluation order;
//   * survivors (typically ~1 per blob) are compacted into a per-warp list of
//     (particle, landmark, blob) triples and evaluated lane-parallel in fp64 exactly as the
//     reference does (both pdfs in the linear domain, so the fp64-underflow match/no-match
//     decision of finding F3 is reproduced, not emulated); the cold part of a landmark
//     (position, covariance blocks, id; 64 B = two DRAM sectors) is fetched only here;
//   * arg-max per (particle, blob) with "first maximum wins" through shared-memory atomics;
//   * the K sequential EKF updates of the group's particles run one (particle, blob) pair per
//     lane; pairs that hit the same landmark of the same particle are ordered in rounds so the
//     second sees the first's result (finding F2);
//   * the weight is the scan-order product of the K factors (:124).
// All arithmetic is fp64; the template parameter is only the landmark STORAGE type.
#include <math.h>

#include "pk_common.cuh"
#include "pk_filter_math.cuh"

namespace pk {

constexpr int kWarpsPerCta = 8;
constexpr int kChunk = 64;      // landmarks per tile
constexpr int kStages = 8;      // tiles in flight per warp
constexpr int kMaxGroup = 8;    // particles per group
constexpr int kMaxItems = 64;   // group * K
constexpr int kListCap = 128;
constexpr unsigned kFull = 0xffffffffu;

struct MeasureArgs {
    double* pose4;
    int* aux2;
    const int* slot;
    unsigned char* pool;
    int* assoc;
    unsigned long long* stats;
    long long M;
    size_t block_bytes;
    int capacity;
    int K;
    int group;  // particles per warp group
    float color_gate_loose;
    pk_params prm;
    double beta[PK_MAX_OBS], cr[PK_MAX_OBS], cg[PK_MAX_OBS], cb[PK_MAX_OBS];
    double dirx[PK_MAX_OBS], diry[PK_MAX_OBS];  // unit((cos b, sin b, 0)) of closest_point :510
    float crf[PK_MAX_OBS], cgf[PK_MAX_OBS], cbf[PK_MAX_OBS];
};

template <typename T>
struct alignas(128) WarpSmem {
    typename Rec<T>::Hot hot[kStages][kChunk];
    double pose[2][kMaxGroup][4];
    unsigned long long best[kMaxItems];
    unsigned long long best_of_j[kMaxItems];
    double factor[kMaxItems];
    int bestj[kMaxItems];
    int list[kListCap];
    int slot_s[2][kMaxGroup];
    int nlive_s[2][kMaxGroup];
    uint64_t bar[kStages];
};

__device__ __forceinline__ int pack_entry(int pl, int k, int j) { return (pl << 26) | (k << 20) | j; }

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
measure_kernel(const __grid_constant__ MeasureArgs A) {
    using Hot = typename Rec<T>::Hot;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem<T>& S = reinterpret_cast<WarpSmem<T>*>(smem_raw)[warp];
    const unsigned lt = lanemask_lt();

    const int K = A.K, GP = A.group, cap = A.capacity;
    const long long M = A.M;
    const long long n_groups = (M + GP - 1) / GP;
    const long long total_warps = (long long)gridDim.x * kWarpsPerCta;
    const long long gw = (long long)blockIdx.x * kWarpsPerCta + warp;
    const long long my_groups = (gw < n_groups) ? (n_groups - gw + total_warps - 1) / total_warps : 0;

    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&S.bar[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    // ---- producer state (warp-uniform) ------------------------------------------------------
    long long p_git = 0, p_tiles = 0;
    int p_pl = 0, p_chunk = 0;
    long long c_git = 0, c_tiles = 0;
    int nx_slot = 0, nx_nlive = 0;  // lane pl holds slot / n_live of particle pl of the next group to open
    auto fetch_info = [&](long long git) {
        nx_slot = 0;
        nx_nlive = 0;
        if (git < my_groups && lane < GP) {
            long long p = (gw + git * total_warps) * GP + lane;
            if (p < M) {
                nx_slot = A.slot[p];
                nx_nlive = A.aux2[2 * p];
            }
        }
    };
    fetch_info(0);

    auto produce = [&]() {
        while (p_git < my_groups && p_git <= c_git + 1 && p_tiles < c_tiles + kStages) {
            const int par = (int)(p_git & 1);
            const long long grp = gw + p_git * total_warps;
            const long long p0 = grp * GP;
            const int gpn = (int)min((long long)GP, M - p0);
            const bool first = (p_pl == 0 && p_chunk == 0);
            if (first) {  // open the group: publish slot / n_live, prefetch the next group's
                if (lane < kMaxGroup) {
                    S.slot_s[par][lane] = nx_slot;
                    S.nlive_s[par][lane] = nx_nlive;
                }
                fetch_info(p_git + 1);
                __syncwarp();
            }
            const int nlive = S.nlive_s[par][p_pl];
            const int nch = max(1, (nlive + kChunk - 1) / kChunk);
            const int nl = max(0, min(kChunk, nlive - p_chunk * kChunk));
            const int stage = (int)(p_tiles % kStages);
            if (lane == 0) {
                const unsigned hot_b = (unsigned)(nl * (int)sizeof(Hot));
                const unsigned pose_b = first ? (unsigned)(gpn * 32) : 0u;
                fence_proxy_async();
                mbar_arrive_expect_tx(&S.bar[stage], hot_b + pose_b);
                if (hot_b)
                    tma_load_1d(&S.hot[stage][0],
                                A.pool + (size_t)S.slot_s[par][p_pl] * A.block_bytes + (size_t)p_chunk * kChunk * sizeof(Hot),
                                hot_b, &S.bar[stage]);
                if (pose_b) tma_load_1d(&S.pose[par][0][0], A.pose4 + 4 * p0, pose_b, &S.bar[stage]);
            }
            ++p_tiles;
            if (++p_chunk >= nch) {
                p_chunk = 0;
                if (++p_pl >= gpn) {
                    p_pl = 0;
                    ++p_git;
                }
            }
        }
    };

    unsigned long long st_matched = 0, st_unmatched = 0, st_eval = 0, st_same = 0, st_promoted = 0;
    unsigned st_flags = 0;

    for (long long git = 0; git < my_groups; ++git) {
        c_git = git;
        const int par = (int)(git & 1);
        const long long grp = gw + git * total_warps;
        const long long p0 = grp * GP;
        const int gpn = (int)min((long long)GP, M - p0);
        const int nitems = gpn * K;
        produce();  // opens this group if it is not open yet
        for (int w = lane; w < kMaxItems; w += 32) {
            S.best[w] = 0ull;
            S.best_of_j[w] = 0ull;
            S.bestj[w] = 0x7fffffff;
        }
        __syncwarp();
        int list_n = 0;

        // Exact evaluation of the listed (particle, landmark, blob) triples + running arg-max.
        auto process_list = [&]() {
            __syncwarp();
            for (int base = 0; base < list_n; base += 32) {
                const int idx = base + lane;
                const bool active = idx < list_n;
                double Lk = 0.0;
                int item = 0, j = 0;
                if (active) {
                    const int e = S.list[idx];
                    const int pl = e >> 26, k = (e >> 20) & 63;
                    j = e & 0xfffff;
                    item = pl * K + k;
                    Landmark L;
                    load_landmark<T>(A.pool + (size_t)S.slot_s[par][pl] * A.block_bytes, cap, j, L);
                    Lk = match_likelihood(L, S.pose[par][pl][0], S.pose[par][pl][1], S.pose[par][pl][2], A.beta[k],
                                          A.cr[k], A.cg[k], A.cb[k], A.dirx[k], A.diry[k], A.prm, st_flags);
                    st_eval += 1;
                }
                // match_one :369-381: strict '>' from 0.0, first maximum (lowest slot) wins
                const bool pos = active && (Lk > 0.0);
                const unsigned long long bits = (unsigned long long)__double_as_longlong(Lk);
                if (pos) atomicMax(&S.best[item], bits);
                __syncwarp();
                const bool win = pos && (S.best[item] == bits);
                if (win && S.best_of_j[item] != bits) S.bestj[item] = 0x7fffffff;  // winner of a smaller value: drop
                __syncwarp();
                if (win) atomicMin(&S.bestj[item], j);
                __syncwarp();
                if (win) S.best_of_j[item] = bits;
                __syncwarp();
            }
            list_n = 0;
        };

        // ---- stream the hot tiles of the group's particles: colour-gate pre-filter ---------------
        for (int pl = 0; pl < gpn; ++pl) {
            const int nlive = S.nlive_s[par][pl];
            const int nch = max(1, (nlive + kChunk - 1) / kChunk);
            for (int ch = 0; ch < nch; ++ch) {
                produce();
                const int stage = (int)(c_tiles % kStages);
                mbar_wait(&S.bar[stage], (unsigned)((c_tiles / kStages) & 1));
                const int nl = max(0, min(kChunk, nlive - ch * kChunk));
                const Hot* hot = &S.hot[stage][0];
                for (int row = 0; row * 32 < nl; ++row) {
                    const int jl = row * 32 + lane;
                    const bool valid = jl < nl;
                    unsigned long long mask = 0ull;
                    if (valid) {
                        const Hot h = hot[jl];
                        if (sizeof(T) == 4) {
                            // fp32 screen with a loose bound; the exact fp64 test is in match_likelihood
                            const float hr = (float)h.r, hg = (float)h.g, hb = (float)h.b;
                            const bool big = !(fmaxf(fmaxf(fabsf(hr), fabsf(hg)), fabsf(hb)) < 4096.0f);
                            for (int k = 0; k < K; ++k) {
                                const float dr = A.crf[k] - hr, dg = A.cgf[k] - hg, db = A.cbf[k] - hb;
                                const float cd = dr * dr + dg * dg + db * db;
                                if (!(cd > A.color_gate_loose) || big) mask |= 1ull << k;
                            }
                        } else {
                            const double hr = (double)h.r, hg = (double)h.g, hb = (double)h.b;
                            for (int k = 0; k < K; ++k) {
                                const double dr = A.cr[k] - hr, dg = A.cg[k] - hg, db = A.cb[k] - hb;
                                const double cd = dr * dr + dg * dg + db * db;
                                if (!(fabs(cd) > A.prm.color_gate)) mask |= 1ull << k;
                            }
                        }
                    }
                    const int j = ch * kChunk + jl;
                    for (;;) {
                        const bool has = mask != 0ull;
                        const unsigned b = __ballot_sync(kFull, has);
                        if (b == 0u) break;
                        const int n = __popc(b);
                        if (list_n + n > kListCap) process_list();
                        if (has) {
                            const int kk = __ffsll((long long)mask) - 1;
                            mask &= mask - 1ull;
                            S.list[list_n + __popc(b & lt)] = pack_entry(pl, kk, j);
                        }
                        list_n += n;
                    }
                }
                __syncwarp();  // every lane is done with this stage before it is refilled
                ++c_tiles;
            }
        }
        process_list();

        // ---- sequential EKF updates (:88-124), one (particle, blob) pair per lane ------------------
        for (int ibase = 0; ibase < nitems; ibase += 32) {
            const int w = ibase + lane;
            const bool act = w < nitems;
            const int pl = act ? w / K : 0, k = act ? w % K : 0;
            int j = -1;
            if (act && S.best[w] != 0ull) j = S.bestj[w];
            const bool matched = act && j >= 0;
            const int key = matched ? (pl * cap + j) : (-1 - lane);
            const unsigned peers = __match_any_sync(kFull, key);
            const int rank = __popc(peers & lt);
            const int maxrank = __reduce_max_sync(kFull, matched ? rank : 0);
            double factor = A.prm.no_match_weight;  // :95 / :851-857
            int id_out = 0;
            for (int r = 0; r <= maxrank; ++r) {
                if (matched && rank == r) {
                    int promoted = 0;
                    factor = ekf_update<T>(A.pool + (size_t)S.slot_s[par][pl] * A.block_bytes, cap, j, S.pose[par][pl][0],
                                           S.pose[par][pl][1], A.beta[k], A.cr[k], A.cg[k], A.cb[k], A.prm, id_out,
                                           st_flags, promoted);
                    st_promoted += promoted;
                    if (r > 0) st_same += 1;
                }
                __syncwarp();
            }
            if (act) {
                A.assoc[(p0 + pl) * K + k] = id_out;
                S.factor[w] = factor;
                S.bestj[w] = id_out;  // reuse as the id list for the orphan count below
                if (matched) st_matched += 1; else st_unmatched += 1;
            }
        }
        __syncwarp();
        if (lane < gpn) {
            // particles[i].weight = 1 (:73); weight *= factor in scan order (:95, :124)
            double wgt = 1.0;
            int orphans = 0;
            for (int k = 0; k < K; ++k) {
                wgt *= S.factor[lane * K + k];
                orphans += (S.bestj[lane * K + k] == 0);
            }
            if (!isfinite(wgt)) st_flags |= PK_FLAG_NONFINITE_WEIGHT;
            A.pose4[4 * (p0 + lane) + 3] = wgt;
            // add_orphaned_reading bumps next_id once per unseen blob (:745-746)
            if (orphans) A.aux2[2 * (p0 + lane) + 1] += orphans;
        }
        __syncwarp();
    }

    // ---- statistics: one atomic per warp per counter ------------------------------------------------
    for (int o = 16; o > 0; o >>= 1) {
        st_matched += __shfl_xor_sync(kFull, st_matched, o);
        st_unmatched += __shfl_xor_sync(kFull, st_unmatched, o);
        st_eval += __shfl_xor_sync(kFull, st_eval, o);
        st_same += __shfl_xor_sync(kFull, st_same, o);
        st_promoted += __shfl_xor_sync(kFull, st_promoted, o);
        st_flags |= __shfl_xor_sync(kFull, st_flags, o);
    }
    if (lane == 0 && A.stats != nullptr) {
        if (st_matched) atomicAdd(&A.stats[PK_STAT_MATCHED], st_matched);
        if (st_unmatched) atomicAdd(&A.stats[PK_STAT_UNMATCHED], st_unmatched);
        if (st_eval) atomicAdd(&A.stats[PK_STAT_EVALUATED], st_eval);
        if (st_same) atomicAdd(&A.stats[PK_STAT_SAME_LANDMARK], st_same);
        if (st_promoted) atomicAdd(&A.stats[PK_STAT_PROMOTED], st_promoted);
        if (st_flags) atomicOr(&A.stats[PK_STAT_FLAGS], (unsigned long long)st_flags);
    }
}

__global__ void reset_weight_kernel(double* __restrict__ pose4, long long M) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) pose4[4 * i + 3] = 1.0;  // cam_cb :73 with an empty scan
}

template <typename T>
static int launch_measure(const MeasureArgs& args, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(WarpSmem<T>) * kWarpsPerCta;
    if (!configured) {
        PK_CUDA(cudaFuncSetAttribute(measure_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int ctas_per_sm = 0;
    PK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, measure_kernel<T>, kWarpsPerCta * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const long long n_groups = (args.M + args.group - 1) / args.group;
    long long grid = (long long)num_sms() * ctas_per_sm;
    const long long need = (n_groups + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    measure_kernel<T><<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(args);
    PK_LAUNCH_CHECK("measure_kernel");
    return PK_OK;
}

}  // namespace pk

using namespace pk;

extern "C" int pk_measurement_update(double* pose4, int* aux2, const int* slot, void* pool, int capacity, int dtype,
                                     long long M, const double* obs_host, int K, const pk_params* params, int* assoc,
                                     unsigned long long* stats, void* stream) {
    PK_CHECK_ARG(pose4 && aux2 && slot && pool, "null state pointer");
    PK_CHECK_ARG(dtype == PK_DTYPE_F32 || dtype == PK_DTYPE_F64, "dtype");
    PK_CHECK_ARG(M >= 0, "M < 0");
    PK_CHECK_ARG(K >= 0 && K <= PK_MAX_OBS, "K must be in [0, PK_MAX_OBS]");
    PK_CHECK_ARG(capacity >= 0 && capacity < (1 << 20), "capacity must be < 2^20");
    PK_CHECK_ARG(params != nullptr, "params is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) return PK_OK;
    if (K == 0) {
        const int threads = 256;
        reset_weight_kernel<<<(unsigned)((M + threads - 1) / threads), threads, 0, st>>>(pose4, M);
        PK_LAUNCH_CHECK("reset_weight_kernel");
        return PK_OK;
    }
    PK_CHECK_ARG(obs_host != nullptr && assoc != nullptr, "obs_host / assoc is NULL");

    static thread_local MeasureArgs args;
    args.pose4 = pose4;
    args.aux2 = aux2;
    args.slot = slot;
    args.pool = (unsigned char*)pool;
    args.assoc = assoc;
    args.stats = stats;
    args.M = M;
    args.block_bytes = block_bytes(capacity, dtype);
    args.capacity = capacity;
    args.K = K;
    int group = 32 / K;
    if (group < 1) group = 1;
    if (group > kMaxGroup) group = kMaxGroup;
    args.group = group;
    args.prm = *params;
    bool obs_big = false;
    for (int k = 0; k < K; ++k) {
        const double beta = obs_host[4 * k + 0];
        args.beta[k] = beta;
        args.cr[k] = obs_host[4 * k + 1];
        args.cg[k] = obs_host[4 * k + 2];
        args.cb[k] = obs_host[4 * k + 3];
        // closest_point :510 / utils.py:69-76: unit((cos b, sin b, 0.0)) = scale(v, 1.0/length)
        volatile double c = cos(beta), s = sin(beta);
        volatile double length = sqrt(c * c + s * s + 0.0 * 0.0);
        volatile double inv = 1.0 / length;
        args.dirx[k] = c * inv;
        args.diry[k] = s * inv;
        args.crf[k] = (float)args.cr[k];
        args.cgf[k] = (float)args.cg[k];
        args.cbf[k] = (float)args.cb[k];
        for (int q = 1; q < 4; ++q)
            if (!(fabs(obs_host[4 * k + q]) < 4096.0)) obs_big = true;
    }
    // fp32 screen: with |colour| < 4096 the fp32 distance is within 0.1% + 0.1 of the fp64 one, so
    // anything the exact gate would accept also passes the loose one.
    args.color_gate_loose = obs_big ? INFINITY : (float)(params->color_gate * 1.001 + 0.25);
    if (dtype == PK_DTYPE_F32) return launch_measure<float>(args, st);
    return launch_measure<double>(args, st);
}
