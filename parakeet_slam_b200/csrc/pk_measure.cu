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
// Shape of the kernel (persistent; one warp owns a *group* of consecutive particles sized so that
// group x K blobs fills the 32 lanes, i.e. every lane is one (particle, blob) ITEM):
//   * the 4-byte colour KEYS of the group's maps are streamed global -> shared by 1-D TMA bulk
//     copies (cp.async.bulk + mbarrier), one copy per particle issued by its own lane, through a
//     per-warp ring of stages that runs ahead of the consumer across groups;
//   * SCREEN: each lane scans its particle's keys against its blob's key with two integer SIMD
//     instructions per pair (vabsdiff4 + dp4a = squared byte distance, keys read 4 at a time with
//     LDS.128) against a bound that provably contains the reference's colour gate (:441) -- the
//     cheapest and most selective of its gates, and probability_of_match is 0 whenever it fails,
//     whatever the evaluation order.  Hits (about one per item) stay in registers; the lane
//     requests the cold record of its first hit at once into its own staging slot in shared
//     memory with per-thread cp.async copies (a scattered, per-lane fetch -- cp.async.bulk takes
//     warp-uniform operands and would serialise over the lanes);
//   * the warp is software-pipelined across groups: it screens group g+1 (and so has that
//     group's records in flight) BEFORE it evaluates group g, so neither the key stream nor the
//     scattered record fetches expose DRAM latency;
//   * EVALUATE: the lane evaluates its candidates in fp64 exactly as the reference does -- both
//     pdfs in the linear domain, so the fp64-underflow match/no-match decision (finding F3) is
//     reproduced, not emulated -- keeps the first maximum (:369-381), then applies the EKF update
//     re-using the bearing it already computed.  Items that hit the same landmark of the same
//     particle are ordered in rounds so the second sees the first's result (finding F2);
//   * the weight is the scan-order product of the K factors (:124).
// All arithmetic is fp64; the template parameter T is only the landmark STORAGE type, R the number
// of items a lane carries (1 when group x K <= 32, 2 for 32 < K <= 64).
#include <math.h>

#include "pk_common.cuh"
#include "pk_filter_math.cuh"

namespace pk {

// 12 warps per SM at 168 registers either way; small CTAs measured best on B200 (config 2, K2 time per frame:
// 12x1 0.865 ms, 6x2 0.828, 4x3 0.808, 3x4 0.813, 2x6 0.804), there is no CTA-wide synchronisation to amortise.
#ifndef PK_MEASURE_WARPS
#define PK_MEASURE_WARPS 2
#endif
#ifndef PK_MEASURE_MINB
#define PK_MEASURE_MINB 6
#endif
constexpr int kWarpsPerCta = PK_MEASURE_WARPS;
constexpr int kChunk = 64;           // keys per particle per stage
constexpr int kKeyStride = kChunk + 4;  // words; +4 keeps the particles' key rows on distinct banks
constexpr int kStages = 2;           // key stages in flight per warp
constexpr int kMaxGroup = 8;         // particles per group
constexpr int kMaxItems = 64;        // group * K
constexpr int kMaxHits = 8;          // hits per item kept (two in registers, the rest in shared memory)
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
    int group;        // particles per warp group
    int key_thr;      // squared byte-distance bound of the colour screen
    int warp_smem;    // bytes of shared memory per warp
    int keys_off;     // offset of the key ring inside a warp's shared memory
    int rec_off;      // offset of the record staging area
    pk_params prm;
    const ObsTable* tab_dev;  // device-resident blob table (pk_measurement_update_dev), else NULL
    ObsTable tab;             // blob table passed by value (host scan)
};

// fixed part of a warp's shared memory; the key ring [kStages][group][kKeyStride] and the record
// staging area [2][32 * R] follow at keys_off / rec_off
template <int R>
struct alignas(128) WarpSmemT {
    static constexpr int kItems = 32 * R;
    double pose[4][kMaxGroup][4];
    double factor[kItems];
    int ids[kItems];
    int bj[kItems];
    int slot_s[4][kMaxGroup];
    int nlive_s[4][kMaxGroup];
    int nsteps_s[4];
    int more_hits[2][kItems][kMaxHits - 2];  // third and later hits of an item (rare)
    uint64_t key_bar[kStages];
};
static_assert(kMaxItems == 64, "R <= 2");

// how many candidates per item get a shared-memory staging slot: two for the 64-byte f32 record, one for f64
#ifndef PK_STAGED_F32
#define PK_STAGED_F32 2
#endif
template <typename T, typename LM = Landmark>
__host__ __device__ constexpr int staged_candidates() {
    return sizeof(LM) == sizeof(LandmarkF) ? PK_STAGED_F32 : (sizeof(typename Rec<T>::Cold) <= 64 ? 2 : 1);
}

// per-item screen result, kept in registers between screen(g) and evaluate(g)
struct Hits {
    int cnt, c0, c1;
};

// ---------------------------------------------------------------------------------------------
// LM selects the arithmetic of the landmark algebra: Landmark (fp64, every instantiation that must reproduce the
// reference) or LandmarkF (fp32 on fp32 storage, PK_DTYPE_ARITH_F32)
#ifndef PK_MEASURE_MINB_F32
#define PK_MEASURE_MINB_F32 8
#endif
template <typename LM> struct ArithOf { using S = double; using Pre = MatchPre; static constexpr int kMinB = PK_MEASURE_MINB; };
template <> struct ArithOf<LandmarkF> { using S = float; using Pre = MatchPreF; static constexpr int kMinB = PK_MEASURE_MINB_F32; };

template <typename T, int R, typename LM>
__global__ void __launch_bounds__(kWarpsPerCta * 32, ArithOf<LM>::kMinB)
measure_kernel(const __grid_constant__ MeasureArgs A) {
    using Cold = typename Rec<T>::Cold;
    using S_t = typename ArithOf<LM>::S;
    using Pre_t = typename ArithOf<LM>::Pre;
    constexpr unsigned kRecBytes = (unsigned)sizeof(Cold);
    constexpr int kStaged = staged_candidates<T, LM>();  // candidates per item prefetched into shared memory
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + (size_t)warp * A.warp_smem;
    using WarpSmem = WarpSmemT<R>;
    WarpSmem& S = *reinterpret_cast<WarpSmem*>(wbase);
    const uint32_t s_base = smem_u32(wbase);
    const uint32_t s_keys = s_base + (uint32_t)A.keys_off;  // [kStages][GP][kKeyStride] words
    const uint32_t s_rec = s_base + (uint32_t)A.rec_off;    // [2][32 * R] Cold
    const uint32_t s_keybar = smem_u32(&S.key_bar[0]);
    const uint32_t s_pose = smem_u32(&S.pose[0][0][0]);
    const unsigned lt = lanemask_lt();

    const int K = A.K, GP = A.group, cap = A.capacity;
    const int key_thr = A.key_thr;
    const long long M = A.M;
    const long long n_groups = (M + GP - 1) / GP;
    const long long total_warps = (long long)gridDim.x * kWarpsPerCta;
    const long long gw = (long long)blockIdx.x * kWarpsPerCta + warp;
    const long long my_groups = (gw < n_groups) ? (n_groups - gw + total_warps - 1) / total_warps : 0;

    const ObsTable* OT = A.tab_dev ? A.tab_dev : &A.tab;
    // this lane's items: item w = r * 32 + lane -> (particle pl, blob k) within a group
    int it_pl[R], it_k[R];
    unsigned it_key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int w = r * 32 + lane;
        it_pl[r] = w / K;
        it_k[r] = w - it_pl[r] * K;
        it_key[r] = (it_pl[r] < GP) ? OT->okey[it_k[r]] : 0u;
    }

    // the blob of this lane's item(s) never changes: keep its values in registers
    S_t ob_beta[R], ob_r[R], ob_g[R], ob_b[R], ob_dx[R], ob_dy[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int k = (it_pl[r] < GP) ? it_k[r] : 0;
        ob_beta[r] = (S_t)OT->beta[k];
        ob_r[r] = (S_t)OT->cr[k];
        ob_g[r] = (S_t)OT->cg[k];
        ob_b[r] = (S_t)OT->cb[k];
        ob_dx[r] = (S_t)OT->dirx[k];
        ob_dy[r] = (S_t)OT->diry[k];
    }

    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&S.key_bar[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    // ---- key producer (warp-uniform state) -----------------------------------------------------
    long long p_git = 0, s_git = 0;
    unsigned p_cnt = 0, s_cnt = 0;  // stages produced / consumed
    int p_step = 0, p_nsteps = 1;
    int nx_slot = 0, nx_nlive = 0;  // lane pl holds slot / n_live of particle pl of the next group to open
    auto fetch_info = [&](long long git) {
        nx_slot = 0;
        nx_nlive = 0;
        if (git < my_groups && lane < GP) {
            const long long p = (gw + git * total_warps) * GP + lane;
            if (p < M) {
                nx_slot = A.slot[p];
                nx_nlive = A.aux2[2 * p];
            }
        }
    };
    fetch_info(0);

    // Group info buffers are indexed git & 3: group g is being evaluated, g+1 screened, g+2 may
    // already be open in the producer.
    auto produce = [&]() {
        while (p_git < my_groups && p_git <= s_git + 1 && p_cnt - s_cnt < (unsigned)kStages) {
            const int gi = (int)(p_git & 3);
            const long long p0 = (gw + p_git * total_warps) * GP;
            const int gpn = (int)min((long long)GP, M - p0);
            if (p_step == 0) {  // open the group: publish slot / n_live, prefetch the next group's
                if (lane < kMaxGroup) {
                    S.slot_s[gi][lane] = nx_slot;
                    S.nlive_s[gi][lane] = nx_nlive;
                }
                const int maxn = __reduce_max_sync(kFull, nx_nlive);
                p_nsteps = max(1, (maxn + kChunk - 1) / kChunk);
                if (lane == 0) S.nsteps_s[gi] = p_nsteps;
                fetch_info(p_git + 1);
                __syncwarp();
            }
            const unsigned stage = p_cnt % kStages;
            unsigned bytes = 0;
            if (lane < gpn) {
                const int nl = max(0, min(kChunk, S.nlive_s[gi][lane] - p_step * kChunk));
                bytes = ((unsigned)nl * 4u + 15u) & ~15u;
            }
            const unsigned total = __reduce_add_sync(kFull, bytes) + (p_step == 0 ? (unsigned)(gpn * 32) : 0u);
            fence_proxy_async();
            const uint32_t bar = s_keybar + stage * 8u;
            if (lane == 0) {
                mbar_arrive_expect_tx_a(bar, total);
                if (p_step == 0) tma_load_1d_a(s_pose + (uint32_t)gi * (kMaxGroup * 32), A.pose4 + 4 * p0, (unsigned)(gpn * 32), bar);
            }
            __syncwarp();
            // every lane moves its own particle's keys.  (cp.async.bulk takes warp-uniform operands, so this
            // compiles to a short waterfall over the <= 8 issuing lanes; issuing all copies from lane 0 in a
            // loop was measured slower.)
            if (bytes)
                tma_load_1d_a(s_keys + ((stage * (unsigned)GP + (unsigned)lane) * kKeyStride) * 4u,
                              A.pool + (size_t)S.slot_s[gi][lane] * A.block_bytes + (size_t)p_step * kChunk * 4, bytes, bar);
            ++p_cnt;
            if (++p_step >= p_nsteps) {
                p_step = 0;
                ++p_git;
            }
        }
    };

    unsigned long long st_matched = 0, st_unmatched = 0, st_eval = 0, st_same = 0, st_promoted = 0;
    unsigned st_flags = 0;

    // ---- screen(g): colour-key screen of group g + record prefetch --------------------------------
    auto screen = [&](long long git, Hits (&H)[R]) {
        s_git = git;
        const int gi = (int)(git & 3), par = (int)(git & 1);
        const long long p0 = (gw + git * total_warps) * GP;
        const int gpn = (int)min((long long)GP, M - p0);
        const int nitems = gpn * K;
        produce();  // opens this group if it is not open yet
#pragma unroll
        for (int r = 0; r < R; ++r) H[r] = Hits{0, -1, -1};
        const int nsteps = S.nsteps_s[gi];
        for (int step = 0; step < nsteps; ++step) {
            produce();
            const unsigned stage = s_cnt % kStages;
            mbar_wait_a(s_keybar + stage * 8u, (s_cnt / kStages) & 1u);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool act = (r * 32 + lane) < nitems;
                const int pl = act ? it_pl[r] : 0;
                const int nl = act ? max(0, min(kChunk, S.nlive_s[gi][pl] - step * kChunk)) : 0;
                const uint32_t kp = s_keys + ((stage * (unsigned)GP + (unsigned)pl) * kKeyStride) * 4u;
                const unsigned mykey = it_key[r];
                // three integer instructions per key: |difference| per byte, dot product accumulated onto
                // -(threshold + 1) (negative <=> inside the bound), and a funnel shift that pushes the sign bit
                // into the hit mask (key i of a 32-key half ends up at bit 31 - i: reversed afterwards)
                const unsigned neg_thr1 = (unsigned)(-(key_thr + 1));
                unsigned lo = 0u, hi = 0u;
#pragma unroll
                for (int q = 0; q < kChunk / 4; ++q) {
                    const int4 v = lds16_a(kp + 16u * q);
                    const unsigned kk[4] = {(unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const unsigned d = __vabsdiffu4(kk[e], mykey);
                        const unsigned sgn = __dp4a(d, d, neg_thr1);
                        if (4 * q + e < 32) lo = __funnelshift_l(sgn, lo, 1); else hi = __funnelshift_l(sgn, hi, 1);
                    }
                }
                unsigned long long m = ((unsigned long long)__brev(hi) << 32) | __brev(lo);
                m &= (nl >= 64) ? ~0ull : ((1ull << nl) - 1ull);  // keys beyond n_live are stale
                while (m) {  // about one hit per item
                    const int j = step * kChunk + __ffsll((long long)m) - 1;
                    m &= m - 1ull;
                    if (H[r].cnt == 0) H[r].c0 = j;
                    else if (H[r].cnt == 1) H[r].c1 = j;
                    else if (H[r].cnt < kMaxHits) S.more_hits[par][r * 32 + lane][H[r].cnt - 2] = j;
                    H[r].cnt += 1;
                }
            }
            __syncwarp();  // every lane is done with this key stage before it is refilled
            ++s_cnt;
        }
        // request the cold record of each item's first hit(s) into the lane's staging slot.  The 32 slots of a
        // round are consecutive in shared memory, so the warp fetches them TOGETHER: instruction i moves 16-byte
        // chunk i*32+lane of that 32-record strip, i.e. the lanes of one instruction cover whole records and every
        // 32-byte sector is requested once.  (Each lane copying its own record 16 bytes at a time asks L2 for every
        // sector twice, from different instructions; the second request missed as well and DRAM read the records
        // twice -- 1.53 GB instead of ~1 GB per launch under ncu.)
        constexpr int kChunksPerRec = (int)(kRecBytes / 16u);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int c = 0; c < kStaged; ++c) {
                const bool have = H[r].cnt > c;
                unsigned long long src = 0ull;
                if (have)
                    src = (unsigned long long)cold_ptr<T>(A.pool + (size_t)S.slot_s[gi][it_pl[r]] * A.block_bytes, cap,
                                                          c == 0 ? H[r].c0 : H[r].c1);
                if (!__any_sync(kFull, have)) continue;
                const uint32_t strip = s_rec + (((unsigned)par * kStaged + (unsigned)c) * 32u * R + (unsigned)(r * 32)) * kRecBytes;
#pragma unroll
                for (int i = 0; i < kChunksPerRec; ++i) {
                    const int chunk = i * 32 + lane;
                    const int rec = chunk / kChunksPerRec, part = chunk - rec * kChunksPerRec;
                    const unsigned long long sp = __shfl_sync(kFull, src, rec);
                    if (sp) cp_async16_a(strip + 16u * (unsigned)chunk, reinterpret_cast<const unsigned char*>(sp) + 16 * part);
                }
            }
        }
        cp_async_commit();
    };

    // ---- evaluate(g): association arg-max + sequential EKF updates + weight ----------------------
    auto evaluate = [&](long long git, const Hits (&H)[R], bool more_in_flight) {
        const int gi = (int)(git & 3), par = (int)(git & 1);
        const long long p0 = (gw + git * total_warps) * GP;
        const int gpn = (int)min((long long)GP, M - p0);
        const int nitems = gpn * K;
        // this group's records were committed one group ago; the next group's may still be in flight
        if (more_in_flight) cp_async_wait<1>(); else cp_async_wait<0>();
        // per-lane association result of the (last) evaluated round, kept in registers
        S_t best_pse = 0;
        int bestj = -1, lastj = -1;
        LM L;
        // ---- association (:84, match_features_to_scan): every blob against the PRE-update map ----
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int w = r * 32 + lane;
            const bool act = w < nitems;
            const int pl = act ? it_pl[r] : 0, k = act ? it_k[r] : 0;
            const unsigned char* block = A.pool + (size_t)S.slot_s[gi][pl] * A.block_bytes;
            const double px = S.pose[gi][pl][0], py = S.pose[gi][pl][1], pth = S.pose[gi][pl][2];
            const int cnt = act ? H[r].cnt : 0;
            // match_one :353-381: arg-max, strict '>' from 0.0, first (lowest slot) maximum wins
            double best = 0.0;
            S_t pse = 0;
            Pre_t pre;
            pre.sure = false;
            best_pse = 0;
            bestj = -1;
            lastj = -1;
            if (cnt > 0) {
                load_staged<T>(s_rec + (((unsigned)par * kStaged) * 32u * R + (unsigned)w) * kRecBytes, L);
                lastj = H[r].c0;
                pre = match_prepare(L, px, py, pth, ob_beta[r], ob_r[r], ob_g[r], ob_b[r], ob_dx[r], ob_dy[r], A.prm,
                                    st_flags);
                pse = pre.pse;
                st_eval += 1;
            }
            // An item with a single colour-compatible landmark only needs the reference's decision
            // `probability > 0.0` (:369-381), and match_prepare can usually PROVE it without the two logs
            // and two exps of the pdf tails (MatchPre::sure).  The value itself is needed only to rank
            // several candidates, or when the product may underflow (finding F3) -- a warp-uniform branch.
#if PK_MATCH_SKIP == 0
            if (cnt > 0) {
                const double Lk = match_finish(pre);
                if (Lk > 0.0) {
                    best = Lk;
                    bestj = lastj;
                    best_pse = pse;
                }
            }
#else
            const bool need_value = cnt > 1 || (cnt == 1 && !pre.sure);
            if (__any_sync(kFull, need_value)) {
                if (cnt > 0) {
                    const double Lk = match_finish(pre);
                    if (Lk > 0.0) {
                        best = Lk;
                        bestj = lastj;
                        best_pse = pse;
                    }
                }
            } else if (cnt > 0) {
                best = 1.0;  // some positive value: never compared against anything
                bestj = lastj;
                best_pse = pse;
            }
#endif
            if (__any_sync(kFull, cnt > 1)) {
                if (cnt <= kMaxHits) {  // further colour-compatible landmarks, in slot order
                    const int maxcnt = __reduce_max_sync(kFull, cnt <= kMaxHits ? cnt : 0);
                    for (int c = 1; c < maxcnt; ++c) {
                        // Most extra hits are false positives of the byte-key screen: apply the exact colour
                        // gate (:441) first and run the full likelihood only if some lane still needs it.
                        bool need = false;
                        if (c < cnt) {
                            const int j = (c == 1) ? H[r].c1 : S.more_hits[par][w][c - 2];
                            if (c < kStaged)
                                load_staged<T>(s_rec + (((unsigned)par * kStaged + (unsigned)c) * 32u * R + (unsigned)w) * kRecBytes, L);
                            else
                                load_landmark<T>(block, cap, j, L);
                            lastj = j;
                            const S_t dr = ob_r[r] - L.r, dg = ob_g[r] - L.g, db = ob_b[r] - L.b;
                            need = !(fabs((double)(dr * dr + dg * dg + db * db)) > A.prm.color_gate);
                        }
                        if (!__any_sync(kFull, need)) continue;
                        if (need) {
                            const double Lk = match_likelihood(L, px, py, pth, ob_beta[r], ob_r[r], ob_g[r], ob_b[r],
                                                               ob_dx[r], ob_dy[r], A.prm, st_flags, pse);
                            st_eval += 1;
                            if (Lk > best) {
                                best = Lk;
                                bestj = lastj;
                                best_pse = pse;
                            }
                        }
                    }
                } else if (cnt > kMaxHits) {
                    // More colour-compatible landmarks than hit registers: walk the particle's keys again
                    // (from global memory this time) and evaluate every landmark that passes the key screen
                    // and the exact colour gate, in slot order.
                    best = 0.0;
                    bestj = -1;
                    lastj = -1;
                    const int nlive = S.nlive_s[gi][pl];
                    const unsigned* gkeys = reinterpret_cast<const unsigned*>(block);
                    const unsigned mykey = it_key[r];
                    for (int j4 = 0; j4 < nlive; j4 += 4) {
                        const int4 kv = ldcg16(gkeys + j4);  // the key region is padded (to 64 bytes)
                        const unsigned kk[4] = {(unsigned)kv.x, (unsigned)kv.y, (unsigned)kv.z, (unsigned)kv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j4 + e;
                            const unsigned dk = __vabsdiffu4(kk[e], mykey);
                            if (j >= nlive || (int)__dp4a(dk, dk, 0u) > key_thr) continue;
                            double cr_, cg_, cb_;
                            load_colour<T>(block, cap, j, cr_, cg_, cb_);
                            const S_t dr = ob_r[r] - (S_t)cr_, dg = ob_g[r] - (S_t)cg_, db = ob_b[r] - (S_t)cb_;
                            if (fabs((double)(dr * dr + dg * dg + db * db)) > A.prm.color_gate) continue;
                            load_landmark<T>(block, cap, j, L);
                            lastj = j;
                            const double Lk = match_likelihood(L, px, py, pth, ob_beta[r], ob_r[r], ob_g[r], ob_b[r],
                                                               ob_dx[r], ob_dy[r], A.prm, st_flags, pse);
                            st_eval += 1;
                            if (Lk > best) {
                                best = Lk;
                                bestj = j;
                                best_pse = pse;
                            }
                        }
                    }
                }
            }
            if (R > 1 && act) {
                S.bj[w] = bestj;
                S.factor[w] = (double)best_pse;
            }
        }
        if (R > 1) __syncwarp();
        // ---- sequential updates (:88-124) in scan order -------------------------------------------
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int w = r * 32 + lane;
            const bool act = w < nitems;
            const int pl = act ? it_pl[r] : 0, k = act ? it_k[r] : 0;
            unsigned char* block = A.pool + (size_t)S.slot_s[gi][pl] * A.block_bytes;
            const double px = S.pose[gi][pl][0], py = S.pose[gi][pl][1];
            if (R > 1) {
                bestj = act ? S.bj[w] : -1;
                best_pse = act ? (S_t)S.factor[w] : (S_t)0;
                lastj = -1;  // records are re-read: an earlier round may have rewritten them
            }
            const bool matched = act && bestj >= 0;
            // items on the same landmark of the same particle go one after the other
            const int key = matched ? (pl * cap + bestj) : (-1 - lane);
            const unsigned peers = __match_any_sync(kFull, key);
            const int rank = __popc(peers & lt);
            const int maxrank = __reduce_max_sync(kFull, matched ? rank : 0);
            double factor = A.prm.no_match_weight;  // :95 / :851-857
            int id_out = 0;
            for (int q = 0; q <= maxrank; ++q) {
                if (matched && rank == q) {
                    // the winner's record is in registers unless another candidate was evaluated after
                    // it, or an earlier blob of this frame has just rewritten the landmark
                    const bool fresh = (q > 0 || lastj != bestj);
                    if (fresh) load_landmark<T>(block, cap, bestj, L);
                    int promoted = 0;
                    bool changed = false;
                    // the bearing computed during association is that of the PRE-update landmark: it can be
                    // re-used only if no earlier blob of this frame has moved the landmark since
                    // L holds the stored colours here, so its key is the one in the hot region: the fp32 instantiation is
                    // memory-latency bound and skips the key store when the update leaves the key alone (-8 % K2 time,
                    // ~0.5 GB less DRAM traffic per launch); the fp64 one is issue bound, where the extra key costs 2 %
                    const unsigned key_before = (sizeof(LM) == sizeof(LandmarkF)) ? stored_key(L, Rec<T>::kDtype) : kNoKey;
                    factor = ekf_update_lm(L, px, py, ob_beta[r], ob_r[r], ob_g[r], ob_b[r], A.prm, id_out, st_flags,
                                           promoted, changed, !fresh, best_pse);
                    if (changed) store_landmark<T>(block, cap, bestj, L, key_before);
                    st_promoted += promoted;
                    if (q > 0) st_same += 1;
                }
                if (maxrank > 0) __syncwarp();
            }
            if (act) {
                A.assoc[(p0 + pl) * K + k] = id_out;
                if (R > 1) {
                    S.factor[w] = factor;
                    S.ids[w] = id_out;
                }
                if (matched) st_matched += 1; else st_unmatched += 1;
            }
            if (R == 1) {
                // particles[i].weight = 1 (:73); weight *= factor in scan order (:95, :124): every lane folds the K
                // factors of its own particle (lanes pl*K .. pl*K+K-1), the particle's first lane stores
                const int base = pl * K;
                double wgt = 1.0;
                for (int k2 = 0; k2 < K; ++k2) wgt *= __shfl_sync(kFull, factor, base + k2);
                const unsigned unseen = __ballot_sync(kFull, act && id_out == 0);
                if (act && k == 0) {
                    if (!isfinite(wgt)) st_flags |= PK_FLAG_NONFINITE_WEIGHT;
                    A.pose4[4 * (p0 + pl) + 3] = wgt;
                    // add_orphaned_reading bumps next_id once per unseen blob (:745-746)
                    const int orphans = __popc((unseen >> base) & (K >= 32 ? 0xffffffffu : ((1u << K) - 1u)));
                    if (orphans) A.aux2[2 * (p0 + pl) + 1] += orphans;
                }
            }
        }
        if (R > 1) {
            __syncwarp();
            if (lane < gpn) {
                double wgt = 1.0;
                int orphans = 0;
                for (int k = 0; k < K; ++k) {
                    wgt *= S.factor[lane * K + k];
                    orphans += (S.ids[lane * K + k] == 0);
                }
                if (!isfinite(wgt)) st_flags |= PK_FLAG_NONFINITE_WEIGHT;
                A.pose4[4 * (p0 + lane) + 3] = wgt;
                if (orphans) A.aux2[2 * (p0 + lane) + 1] += orphans;
            }
        }
        __syncwarp();
    };

    // ---- software pipeline over this warp's groups -----------------------------------------------
    Hits Hcur[R], Hnext[R];
#pragma unroll
    for (int r = 0; r < R; ++r) Hcur[r] = Hnext[r] = Hits{0, -1, -1};
    for (long long git = -1; git < my_groups; ++git) {
        if (git + 1 < my_groups) screen(git + 1, Hnext);
        if (git >= 0) evaluate(git, Hcur, git + 1 < my_groups);
#pragma unroll
        for (int r = 0; r < R; ++r) Hcur[r] = Hnext[r];
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

template <typename T, int R, typename LM>
static int launch_measure(MeasureArgs& args, cudaStream_t st) {
    static int configured_smem = -1;
    auto align128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    args.keys_off = (int)align128(sizeof(WarpSmemT<R>));
    args.rec_off = (int)align128(args.keys_off + (size_t)kStages * args.group * kKeyStride * 4);
    args.warp_smem = (int)align128(args.rec_off + (size_t)2 * staged_candidates<T, LM>() * 32 * R * sizeof(typename Rec<T>::Cold));
    const size_t smem = (size_t)args.warp_smem * kWarpsPerCta;
    if ((int)smem > configured_smem) {
        PK_CUDA(cudaFuncSetAttribute(measure_kernel<T, R, LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_smem = (int)smem;
    }
    int ctas_per_sm = 0;
    PK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, measure_kernel<T, R, LM>, kWarpsPerCta * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const long long n_groups = (args.M + args.group - 1) / args.group;
    long long grid = (long long)num_sms() * ctas_per_sm;
    const long long need = (n_groups + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    measure_kernel<T, R, LM><<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(args);
    PK_LAUNCH_CHECK("measure_kernel");
    return PK_OK;
}

template <typename T, typename LM>
static int dispatch_measure(MeasureArgs& args, cudaStream_t st) {
    if (args.K <= 32) return launch_measure<T, 1, LM>(args, st);
    return launch_measure<T, 2, LM>(args, st);
}

}  // namespace pk

using namespace pk;

// blob table from a device-resident scan: one thread per blob, the arithmetic of the host path
// (closest_point :510 / utils.py:69-76 with separate roundings; cos/sin are CUDA's, <= 2 ulp from libm)
__global__ void obs_table_kernel(const double* __restrict__ obs, int K, pk::ObsTable* __restrict__ tab) {
    const int k = threadIdx.x;
    if (k >= PK_MAX_OBS) return;
    if (k >= K) {
        tab->okey[k] = 0u;
        return;
    }
    const double beta = obs[4 * k], r = obs[4 * k + 1], g = obs[4 * k + 2], b = obs[4 * k + 3];
    double s, c;
    sincos(beta, &s, &c);
    const double length = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(c, c), __dmul_rn(s, s)), 0.0));
    const double inv = __ddiv_rn(1.0, length);
    tab->beta[k] = beta;
    tab->cr[k] = r;
    tab->cg[k] = g;
    tab->cb[k] = b;
    tab->dirx[k] = __dmul_rn(c, inv);
    tab->diry[k] = __dmul_rn(s, inv);
    tab->okey[k] = pk::color_key(r, g, b);
}

static int measurement_common(double* pose4, int* aux2, const int* slot, void* pool, int capacity, int dtype, long long M,
                              const double* obs_host, const double* obs_dev, void* table_ws, int K,
                              const pk_params* params, int* assoc, unsigned long long* stats, void* stream) {
    PK_CHECK_ARG(pose4 && aux2 && slot && pool, "null state pointer");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
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
    PK_CHECK_ARG((obs_host != nullptr || obs_dev != nullptr) && assoc != nullptr, "obs / assoc is NULL");
    if (stats != nullptr) PK_CUDA(cudaMemsetAsync(stats, 0, PK_NUM_STATS * sizeof(unsigned long long), st));

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
    args.tab_dev = nullptr;
    if (obs_dev != nullptr) {
        PK_CHECK_ARG(table_ws != nullptr, "table workspace is NULL");
        obs_table_kernel<<<1, PK_MAX_OBS, 0, st>>>(obs_dev, K, (ObsTable*)table_ws);
        PK_LAUNCH_CHECK("obs_table_kernel");
        args.tab_dev = (const ObsTable*)table_ws;
    } else {
        auto key_of = [](double c) -> unsigned {
            if (!(c == c)) return 0u;  // NaN
            c = c < 0.0 ? 0.0 : (c > 255.0 ? 255.0 : c);
            return (unsigned)nearbyint(c);
        };
        ObsTable& T = args.tab;
        for (int k = 0; k < PK_MAX_OBS; ++k) T.okey[k] = 0u;
        for (int k = 0; k < K; ++k) {
            const double beta = obs_host[4 * k + 0];
            T.beta[k] = beta;
            T.cr[k] = obs_host[4 * k + 1];
            T.cg[k] = obs_host[4 * k + 2];
            T.cb[k] = obs_host[4 * k + 3];
            // closest_point :510 / utils.py:69-76: unit((cos b, sin b, 0.0)) = scale(v, 1.0/length)
            volatile double c = cos(beta), s = sin(beta);
            volatile double length = sqrt(c * c + s * s + 0.0 * 0.0);
            volatile double inv = 1.0 / length;
            T.dirx[k] = c * inv;
            T.diry[k] = s * inv;
            T.okey[k] = key_of(T.cr[k]) | (key_of(T.cg[k]) << 8) | (key_of(T.cb[k]) << 16);
        }
    }
    // Colour screen bound (DESIGN.md "colour keys").  Keys are the colours clamped to [0,255] and
    // rounded, so per channel |key difference| <= |true difference| + 1.  If the exact gate accepts
    // (sum d^2 <= gate) then sum |d| <= sqrt(3*gate) and the squared key distance is at most
    // gate + 2*sqrt(3*gate) + 3.  Anything above that bound cannot pass the reference's gate.
    const double g = params->color_gate;
    double bound = -1.0;  // negative gate: nothing passes
    if (g >= 0.0) bound = floor(g + 2.0 * sqrt(3.0 * g) + 3.0) + 1.0;
    if (!(g == g)) bound = 2.0e9;  // NaN gate: `abs(cd) > nan` is False, everything passes
    args.key_thr = (int)(bound > 2.0e9 ? 2.0e9 : bound);
    if (dtype_arith_f32(dtype)) return dispatch_measure<float, LandmarkF>(args, st);
    if (dtype_base(dtype) == PK_DTYPE_F32) return dispatch_measure<float, Landmark>(args, st);
    return dispatch_measure<double, Landmark>(args, st);
}

extern "C" int pk_measurement_update(double* pose4, int* aux2, const int* slot, void* pool, int capacity, int dtype,
                                     long long M, const double* obs_host, int K, const pk_params* params, int* assoc,
                                     unsigned long long* stats, void* stream) {
    PK_CHECK_ARG(K == 0 || M == 0 || obs_host != nullptr, "obs_host is NULL");
    return measurement_common(pose4, aux2, slot, pool, capacity, dtype, M, obs_host, nullptr, nullptr, K, params, assoc,
                              stats, stream);
}

extern "C" long long pk_obs_table_bytes(void) { return (long long)sizeof(ObsTable); }

extern "C" int pk_measurement_update_dev(double* pose4, int* aux2, const int* slot, void* pool, int capacity, int dtype,
                                         long long M, const double* obs_dev, int K, const pk_params* params, int* assoc,
                                         unsigned long long* stats, void* table_ws, void* stream) {
    PK_CHECK_ARG(K == 0 || M == 0 || obs_dev != nullptr, "obs_dev is NULL");
    return measurement_common(pose4, aux2, slot, pool, capacity, dtype, M, nullptr, obs_dev, table_ws, K, params, assoc,
                              stats, stream);
}
