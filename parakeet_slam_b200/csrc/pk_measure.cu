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
//   * the 4-byte colour KEYS of the group's maps (and its pose records) are streamed global -> shared with per-thread
//     16-byte cp.async copies (16 lanes cover a particle's 64 keys, i.e. whole 128-byte lines per instruction) through a
//     per-warp ring of stages that runs ahead of the consumer across groups;
//   * SCREEN: each lane scans its particle's keys against its blob's key with three integer instructions per pair
//     (vabsdiff4 + dp4a = squared byte distance accumulated onto -(bound + 1), funnel shift of the sign into the hit
//     mask; keys read 4 at a time with LDS.128) against a bound that provably contains the reference's colour gate
//     (:441) -- the cheapest and most selective of its gates, and probability_of_match is 0 whenever it fails,
//     whatever the evaluation order.  Hits (about one per item) stay in registers; the cold records of the items'
//     first hits are requested into a 32-slot strip of shared memory by the whole warp together (per-thread cp.async
//     copies whose lanes cover whole records), a second hit by its own lane, third and later ones through the
//     candidate ring during the evaluation;
//   * the warp is software-pipelined across groups: it screens group g+1 (and so has that
//     group's records in flight) BEFORE it evaluates group g, so neither the key stream nor the
//     scattered record fetches expose DRAM latency;
//   * EVALUATE: the lane evaluates its candidates in fp64 exactly as the reference does -- both
//     pdfs in the linear domain, so the fp64-underflow match/no-match decision (finding F3) is
//     reproduced, not emulated -- keeps the first maximum (:369-381), then applies the EKF update
//     re-using the bearing it already computed.  Items that hit the same landmark of the same
//     particle are ordered in rounds so the second sees the first's result (finding F2);
//   * the weight is the scan-order product of the K factors (:124).
// Template parameters: T the landmark STORAGE type, R the number of items a lane carries (1 when group x K <= 32, 2 for
// 32 < K <= 64), LM the landmark ALGEBRA (Landmark: fp64, every instantiation that must reproduce the reference;
// LandmarkF: fp32 on fp32 records -- poses, differences pose - landmark, weights stay fp64).
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
#ifndef PK_MAX_HITS
#define PK_MAX_HITS 16
#endif
constexpr int kMaxHits = PK_MAX_HITS;  // hits per item kept (two in registers, the rest in shared memory)
constexpr unsigned kFull = 0xffffffffu;

struct MeasureArgs {
    double* pose4;
    int* aux2;
    const int* slot;
    unsigned char* pool;
    int* assoc;
    unsigned long long* stats;
    int M;            // particles of this launch (< 2^31: all group / particle indices are 32-bit)
    unsigned block_bytes;  // < 2^32 (capacity < 2^16): slot * block_bytes is one 32 x 32 -> 64-bit multiply-add
    unsigned hot_bytes;  // hot_region_bytes(capacity): offset of the cold region inside a block
    int capacity;
    int K;
    int group;        // particles per warp group
    int key_thr;      // squared byte-distance bound of the colour screen
    int key_thr1;     // the same bound as a sum of absolute byte differences (contains the squared one)
    double log_no_match;  // log(prm.no_match_weight) (PK_MODEL_LOG_WEIGHTS)
    int warp_smem;    // bytes of shared memory per warp
    int keys_off;     // offset of the key ring inside a warp's shared memory
    int rec_off;      // offset of the record staging area
    pk_params prm;
    const ObsTable* tab_dev;  // device-resident blob table (pk_measurement_update_dev), else NULL
    ObsTable tab;             // blob table passed by value (host scan)
};

// fixed part of a warp's shared memory; the key ring [kStages][group][kKeyStride] and the record
// staging area [strips][32 * R] follow at keys_off / rec_off
template <int R>
struct alignas(128) WarpSmemT {
    static constexpr int kItems = 32 * R;
    double pose[4][kMaxGroup][4];
    double factor[kItems];
    int slot_s[4][kMaxGroup];
    int nlive_s[4][kMaxGroup];
    unsigned short more_hits[2][kItems][kMaxHits - 2];  // third and later hits of an item (rare)
    int unseen_acc;  // R > 1: unseen blobs of the particle, summed over the rounds
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
// Third and later candidates of an item (colour-ambiguous maps, duplicate landmarks of the spawning pipeline): their
// records are requested two candidates ahead with per-lane cp.async copies into a ring of three strips -- candidate c
// lives in strip c % 3, entry = item -- instead of being loaded synchronously when their turn comes, so the exact
// evaluations of candidates c and c + 1 hide the DRAM latency of candidate c + 2.  (64-byte records only.)
#ifndef PK_CLOOP_PREFETCH
#define PK_CLOOP_PREFETCH 1
#endif
template <typename T, typename LM = Landmark>
__host__ __device__ constexpr bool candidate_ring() {
    return PK_CLOOP_PREFETCH != 0 && staged_candidates<T, LM>() == 2;
}
template <typename T, typename LM = Landmark>
__host__ __device__ constexpr int staging_strips() {
    return candidate_ring<T, LM>() ? 3 : staged_candidates<T, LM>();
}

// per-item screen result, kept in registers between screen(g) and evaluate(g)
struct Hits {
    int cnt, c0, c1;
};

// ---------------------------------------------------------------------------------------------
// LM selects the arithmetic of the landmark algebra: Landmark (fp64, every instantiation that must reproduce the
// reference) or LandmarkF (fp32 on fp32 storage, PK_DTYPE_ARITH_F32)
// record prefetch of an item's first hit: 1 = the warp fetches the 32-record strip together, 0 = every lane its own
// colour screen: 0 = squared byte distance (three instructions per key), 1 = sum of absolute byte differences (two:
// VABSDIFF4.U8.ACC + funnel shift).  Measured (tools/k2_variants.py): at 64 landmarks both cost the same (K2 0.438 ms
// either way: the kernel is not bound by its instruction count), on maps of hundreds of landmarks the L1 form's 1.65x
// false positives push items over the hit list and into the key re-walk (config 5: 57.1 s instead of 49.2 s).
#ifndef PK_SCREEN_SAD
#define PK_SCREEN_SAD 0
#endif
#ifndef PK_COOP_PREFETCH
#define PK_COOP_PREFETCH 1
#endif
#ifndef PK_MEASURE_MINB_F32
#define PK_MEASURE_MINB_F32 8
#endif
// PK_K2_REGCAP_F32 > 0: cap the fp32-algebra kernel's registers directly (__maxnreg__) instead of through the
// minimum-CTAs bound; the occupancy query at launch then decides how many CTAs are resident
#ifndef PK_K2_REGCAP_F32
#define PK_K2_REGCAP_F32 0
#endif
template <typename LM> struct ArithOf { using S = double; using Pre = MatchPre; static constexpr int kMinB = PK_MEASURE_MINB; static constexpr int kMaxReg = 168; };
template <> struct ArithOf<LandmarkF> { using S = float; using Pre = MatchPreF; static constexpr int kMinB = PK_MEASURE_MINB_F32; static constexpr int kMaxReg = PK_K2_REGCAP_F32 > 0 ? PK_K2_REGCAP_F32 : 128; };

template <typename T, int R, typename LM>
#if PK_K2_REGCAP_F32 > 0
__global__ void __maxnreg__(ArithOf<LM>::kMaxReg)
#else
__global__ void __launch_bounds__(kWarpsPerCta * 32, ArithOf<LM>::kMinB)
#endif
measure_kernel(const __grid_constant__ MeasureArgs A) {
    using Cold = typename Rec<T>::Cold;
    using S_t = typename ArithOf<LM>::S;
    using Pre_t = typename ArithOf<LM>::Pre;
    constexpr unsigned kRecBytes = (unsigned)sizeof(Cold);
    constexpr int kStaged = staged_candidates<T, LM>();  // candidates per item prefetched into shared memory
    constexpr bool kRing = candidate_ring<T, LM>();      // later candidates travel through a ring of three strips
    constexpr int kChunksPerRec = (int)(kRecBytes / 16u);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + (size_t)warp * A.warp_smem;
    using WarpSmem = WarpSmemT<R>;
    WarpSmem& S = *reinterpret_cast<WarpSmem*>(wbase);
    const uint32_t s_base = smem_u32(wbase);
    const uint32_t s_keys = s_base + (uint32_t)A.keys_off;  // [kStages][GP][kKeyStride] words
    const uint32_t s_rec = s_base + (uint32_t)A.rec_off;    // [staging_strips][32 * R] Cold
    const uint32_t s_pose = smem_u32(&S.pose[0][0][0]);
    const unsigned lt = lanemask_lt();

    const int K = A.K, GP = A.group, cap = A.capacity;
    const int key_thr = A.key_thr;
    const int M = A.M;
    const int n_groups = (M + GP - 1) / GP;
    const int total_warps = (int)gridDim.x * kWarpsPerCta;
    const int gw = (int)blockIdx.x * kWarpsPerCta + warp;
    const int my_groups = (gw < n_groups) ? (n_groups - gw + total_warps - 1) / total_warps : 0;
    const unsigned hot = A.hot_bytes;
    const unsigned bb = A.block_bytes;
    // byte offset of a landmark block inside the pool (slots are non-negative)
    auto blk_off = [bb](int sl) -> size_t { return (size_t)((unsigned long long)(unsigned)sl * (unsigned long long)bb); };

    const ObsTable* OT = A.tab_dev ? A.tab_dev : &A.tab;
    // two blobs of this frame may hit the same landmark only if their colours are close (ObsTable::twins); when no
    // pair is, the same-landmark ordering of the update phase (finding F2) is skipped altogether
    const bool twins = OT->twins != 0u;
    // PK_MODEL_LOG_WEIGHTS: importance factors and particle weights are carried as logarithms
    const bool log_w = (A.prm.model & PK_MODEL_LOG_WEIGHTS) != 0;
    const double log_no_match = A.log_no_match;  // computed by the host: nothing of an fp64 log inside the main loop
    // this lane's items: item w = r * 32 + lane -> (particle pl, blob k) within a group
    int it_pl[R], it_k[R];
    unsigned it_key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int w = r * 32 + lane;
        it_pl[r] = w / K;
        it_k[r] = w - it_pl[r] * K;
        it_key[r] = (it_pl[r] < GP) ? OT->okey[it_k[r]] : 0u;
        if (it_pl[r] >= GP) it_pl[r] = GP;  // inactive lane (GP * K < 32)
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

    // ---- key producer (warp-uniform state) -----------------------------------------------------
    // A flat sequence of steps (group, 64-key chunk) goes through the ring of kStages key stages; produce_one()
    // issues the next step and is called once per consumed step, so the producer stays kStages - 1 steps ahead.
    // Keys and poses move with per-thread 16-byte cp.async copies (LDGSTS): a lane owns one 16-byte piece, 16 lanes
    // cover a particle's 64 keys, so one instruction requests whole 128-byte lines -- no mbarrier, no proxy fence,
    // no per-lane waterfall over uniform bulk-copy operands (this replaced 1-D cp.async.bulk copies; a warp-private
    // stream of 1 KB per step gains nothing from the bulk engine).  Every call commits exactly ONE cp.async group
    // (empty when there is nothing left), and so does prefetch(): the wait_group counts below rely on that.
    // Group info buffers are indexed it & 3: group it is being evaluated, it + 1 screened, it + 2 may be open.
    int p_it = 0, p_step = 0, p_ns = 1;
    unsigned p_cnt = 0, s_cnt = 0;  // steps produced / consumed
    int nx_slot = 0, nx_nlive = 0;  // lane pl: slot / n_live of particle pl of the next group to open
    int op_slot = 0, op_nlive = 0;  // ... of the group that is open in the producer
    // (loaded unconditionally from a clamped index so that the values land in their registers without a
    // select that would wait for them; lanes that own no particle are zeroed when the group is opened)
    auto fetch_info = [&](int it) {
        const int p = min((gw + min(it, max(my_groups - 1, 0)) * total_warps) * GP + lane, M - 1);
        nx_slot = A.slot[p];
        nx_nlive = A.aux2[2 * p];
    };
    fetch_info(0);
    // key copies: this lane's piece of a particle's 64-key chunk
    const int kc_ch = lane & 15, kc_pl = lane >> 4;
    const uint32_t kc_dst = s_keys + (unsigned)kc_pl * (kKeyStride * 4u) + 16u * (unsigned)kc_ch;
    const unsigned char* kc_src = A.pool + 16 * kc_ch;
    auto produce_one = [&]() {
        if (p_it < my_groups) {
            const int gi = p_it & 3;
            const int p0 = (gw + p_it * total_warps) * GP;
            const int gpn = min(GP, M - p0);
            if (p_step == 0) {  // open the group: publish slot / n_live, fetch its poses, prefetch the next group's info
                op_slot = lane < gpn ? nx_slot : 0;
                op_nlive = lane < gpn ? nx_nlive : 0;
                if (lane < kMaxGroup) {
                    S.slot_s[gi][lane] = op_slot;
                    S.nlive_s[gi][lane] = op_nlive;
                }
                p_ns = max(1, (__reduce_max_sync(kFull, op_nlive) + kChunk - 1) / kChunk);
                fetch_info(p_it + 1);
                if (lane < 2 * gpn)
                    cp_async16_line(s_pose + (uint32_t)gi * (kMaxGroup * 32) + 16u * (unsigned)lane,
                                    reinterpret_cast<const unsigned char*>(A.pose4 + 4 * (size_t)p0) + 16 * lane);
            }
            const unsigned stage = p_cnt % kStages;
            // lane -> 16-byte piece kc_ch of particle kc_pl (+2 per round): two shuffles, one 64-bit multiply-add
            // and one copy per round
            const uint32_t dst0 = kc_dst + stage * (unsigned)GP * (kKeyStride * 4u);
            const int first_key = p_step * kChunk + 4 * kc_ch;  // first key of this lane's piece
            for (int pl0 = 0; pl0 < GP; pl0 += 2) {
                const int pl = min(pl0 + kc_pl, GP - 1);
                const int nl = __shfl_sync(kFull, op_nlive, pl);
                const int sl = __shfl_sync(kFull, op_slot, pl);
                if (pl0 + kc_pl < GP && first_key < nl)
                    cp_async16_line(dst0 + (unsigned)pl0 * (kKeyStride * 4u), kc_src + blk_off(sl) + (unsigned)(p_step * kChunk * 4));
            }
            ++p_cnt;
            if (++p_step >= p_ns) {
                p_step = 0;
                ++p_it;
            }
        }
        cp_async_commit();
    };

    // statistics: warp-uniform 32-bit counts (one warp sees < 2^31 items), folded into the 64-bit totals at the end
    unsigned st_matched = 0, st_unmatched = 0, st_eval = 0, st_same = 0, st_promoted = 0;
    unsigned st_flags = 0;

    // ---- scan(it): colour-key screen of group it ---------------------------------------------------
    auto scan = [&](int it, Hits (&H)[R]) {
        const int gi = it & 3, par = it & 1;
        const int p0 = (gw + it * total_warps) * GP;
        const int nitems = min(GP, M - p0) * K;
#pragma unroll
        for (int r = 0; r < R; ++r) H[r] = Hits{0, -1, -1};
        int my_nl[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool act = (r * 32 + lane) < nitems;
            my_nl[r] = act ? S.nlive_s[gi][it_pl[r]] : 0;
        }
        int maxnl = my_nl[0];
#pragma unroll
        for (int r = 1; r < R; ++r) maxnl = max(maxnl, my_nl[r]);
        const int nsteps = max(1, (__reduce_max_sync(kFull, maxnl) + kChunk - 1) / kChunk);
        for (int step = 0; step < nsteps; ++step) {
            const unsigned stage = s_cnt % kStages;
            // this step's keys: one newer group (the previous group's records) is in flight at step 0, none later
            if (step == 0) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int pl = min(it_pl[r], GP - 1);
                const int nl = my_nl[r] - step * kChunk;  // keys of this chunk that are live (<= 0: none)
                const uint32_t kp = s_keys + ((stage * (unsigned)GP + (unsigned)pl) * kKeyStride) * 4u;
                const unsigned mykey = it_key[r];
#if PK_SCREEN_SAD
                // two integer instructions per key: the sum of the absolute byte differences accumulated onto
                // -(bound + 1) in ONE instruction (VABSDIFF4.U8.ACC; negative <=> inside the bound) and a funnel
                // shift that pushes the sign bit into the hit mask (key i of a 32-key half ends up at bit 31 - i:
                // reversed afterwards).  The L1 bound floor(sqrt(3 * key_thr)) contains the squared-distance bound
                // (Cauchy-Schwarz over the three channels), which contains the reference's gate; its extra false
                // positives (0.18 instead of 0.11 per item at 64 random colours) fail the exact gate later.
                const unsigned neg_thr1 = (unsigned)(-(A.key_thr1 + 1));
#else
                // three integer instructions per key: |difference| per byte, dot product accumulated onto
                // -(threshold + 1) (negative <=> inside the bound), and a funnel shift that pushes the sign bit
                // into the hit mask (key i of a 32-key half ends up at bit 31 - i: reversed afterwards)
                const unsigned neg_thr1 = (unsigned)(-(key_thr + 1));
#endif
                unsigned lo = 0u, hi = 0u;
#pragma unroll
                for (int q = 0; q < kChunk / 4; ++q) {
                    const int4 v = lds16_a(kp + 16u * q);
                    const unsigned kk[4] = {(unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
#if PK_SCREEN_SAD
                        const unsigned sgn = vsad4_acc(kk[e], mykey, neg_thr1);
#else
                        const unsigned d = __vabsdiffu4(kk[e], mykey);
                        const unsigned sgn = __dp4a(d, d, neg_thr1);
#endif
                        if (4 * q + e < 32) lo = __funnelshift_l(sgn, lo, 1); else hi = __funnelshift_l(sgn, hi, 1);
                    }
                }
                // keys beyond n_live are stale: mask them (nl <= 0 clears everything)
                const unsigned vlo = nl >= 32 ? 0xffffffffu : (nl > 0 ? (1u << nl) - 1u : 0u);
                const unsigned vhi = nl >= 64 ? 0xffffffffu : (nl > 32 ? (1u << (nl - 32)) - 1u : 0u);
                const unsigned mlo = __brev(lo) & vlo, mhi = __brev(hi) & vhi;
                // about one hit per item: the first two are extracted without branches, the rest in a rare loop
                const int n = __popc(mlo) + __popc(mhi);
                const int base = step * kChunk;
                const int j0 = base + (mlo ? __ffs((int)mlo) - 1 : 31 + __ffs((int)mhi));
                const unsigned mlo2 = mlo & (mlo - 1u), mhi2 = mlo ? mhi : (mhi & (mhi - 1u));
                const int j1 = base + (mlo2 ? __ffs((int)mlo2) - 1 : 31 + __ffs((int)mhi2));
                const int c_before = H[r].cnt;
                if (step == 0) {  // warp-uniform; the only step of maps up to 64 landmarks
                    H[r].c0 = n > 0 ? j0 : -1;
                    H[r].c1 = n > 1 ? j1 : -1;
                } else {  // selects, not branches: c_before differs from lane to lane
                    H[r].c1 = (c_before == 0 && n > 1) ? j1 : ((c_before == 1 && n > 0) ? j0 : H[r].c1);
                    H[r].c0 = (c_before == 0 && n > 0) ? j0 : H[r].c0;
                }
                H[r].cnt = c_before + n;
                if (__any_sync(kFull, c_before + n > 2)) {
                    unsigned ma = mlo, mb = mhi;
                    int idx = c_before;
                    while (ma | mb) {
                        const int j = base + (ma ? __ffs((int)ma) - 1 : 31 + __ffs((int)mb));
                        if (ma) ma &= ma - 1u; else mb &= mb - 1u;
                        if (idx >= 2 && idx < kMaxHits) S.more_hits[par][r * 32 + lane][idx - 2] = (unsigned short)j;
                        ++idx;
                    }
                }
            }
            __syncwarp();  // every lane is done with this key stage before it is refilled
            ++s_cnt;
            produce_one();
        }
    };

    // ---- prefetch(it): request the cold records of group it's first hits ------------------------------
    // The staging area is single-buffered: this runs after the association phase of the previous group has
    // consumed its records (evaluate calls it between its two phases), a whole update phase and key scan
    // before the records are needed.
    auto prefetch = [&](int it, const Hits (&H)[R]) {
        const int gi = it & 3;
        const int p0 = (gw + it * total_warps) * GP;
        const int nitems = min(GP, M - p0) * K;
        int my_slot[R];
#pragma unroll
        for (int r = 0; r < R; ++r) my_slot[r] = ((r * 32 + lane) < nitems) ? S.slot_s[gi][it_pl[r]] : 0;
        __syncwarp();  // every lane has read its records of the previous group out of the staging area
        // Request the cold record of each item's first hit into the lane's staging slot.  The 32 slots of a round
        // are consecutive in shared memory, so the warp fetches them TOGETHER: instruction i moves 16-byte chunk
        // i*32+lane of that 32-record strip, i.e. the lanes of one instruction cover whole records (each 32-byte
        // sector is requested once and an instruction touches 8 lines instead of 32).  The record's position
        // travels between lanes as ONE 32-bit word: its offset inside the pool in units of 32 bytes.
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const size_t rec0 = blk_off(my_slot[r]) + hot;
#if PK_COOP_PREFETCH
            {
                const bool have = H[r].cnt > 0;
                const unsigned off32 = have ? (unsigned)((rec0 + (size_t)H[r].c0 * kRecBytes) >> 5) : 0xffffffffu;
                const uint32_t strip = s_rec + (unsigned)(r * 32) * kRecBytes;
#pragma unroll
                for (int i = 0; i < kChunksPerRec; ++i) {
                    const int chunk = i * 32 + lane;
                    const int rec = chunk / kChunksPerRec, part = chunk - rec * kChunksPerRec;
                    const unsigned so = __shfl_sync(kFull, off32, rec);
                    if (so != 0xffffffffu) cp_async16_a(strip + 16u * (unsigned)chunk, A.pool + ((size_t)so << 5) + 16 * part);
                }
            }
#else
            if (H[r].cnt > 0) {  // every lane fetches its own record (fewer instructions, four times the L1 wavefronts)
                const unsigned char* src = A.pool + rec0 + (size_t)H[r].c0 * kRecBytes;
                const uint32_t dst = s_rec + (unsigned)(r * 32 + lane) * kRecBytes;
#pragma unroll
                for (int i = 0; i < kChunksPerRec; ++i) cp_async16_a(dst + 16u * i, src + 16 * i);
            }
#endif
            // second hit (one item in ten): fetched by its own lane
            if (kStaged > 1 && H[r].cnt > 1) {
                const unsigned char* src = A.pool + rec0 + (size_t)H[r].c1 * kRecBytes;
                const uint32_t dst = s_rec + (32u * R + (unsigned)(r * 32 + lane)) * kRecBytes;
#pragma unroll
                for (int i = 0; i < kChunksPerRec; ++i) cp_async16_a(dst + 16u * i, src + 16 * i);
            }
        }
        cp_async_commit();
    };

    // ---- evaluate(it): association arg-max + sequential EKF updates + weight ----------------------
    // `Hn` are the hits of the next group (already scanned): its records are requested between the two phases.
    auto evaluate = [&](int it, const Hits (&H)[R], const Hits (&Hn)[R], bool have_next) {
        const int gi = it & 3, par = it & 1;
        const int p0 = (gw + it * total_warps) * GP;
        const int gpn = min(GP, M - p0);
        const int nitems = gpn * K;
        cp_async_wait<1>();  // all but the newest group (the keys requested at the end of the last scan)
        __syncwarp();  // the first-hit strip was filled cooperatively: other lanes' copies must have landed too
        // association result per item: winner slot and its bearing.  With one item per lane (R == 1) and fp32 algebra
        // L ends up holding the winner's PRE-update record; otherwise L_j says which record L holds.
        constexpr bool kKeepWinner = (R == 1) && (sizeof(LM) == sizeof(LandmarkF));
        S_t best_pse[R];
        int bestj[R];
        LM L;
        int L_j = -1;
        // ---- association (:84, match_features_to_scan): every blob against the PRE-update map ----
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int w = r * 32 + lane;
            const bool act = w < nitems;
            const int pl = act ? it_pl[r] : 0;
            const unsigned char* block = A.pool + blk_off(S.slot_s[gi][pl]);
            const double px = S.pose[gi][pl][0], py = S.pose[gi][pl][1], pth = S.pose[gi][pl][2];
            const int cnt = act ? H[r].cnt : 0;
            // match_one :353-381: arg-max, strict '>' from 0.0, first (lowest slot) maximum wins
            Pre_t pre;
            pre.sure = false;
            // likelihood (fp64 algebra) or log-domain rank (fp32 algebra) of the best landmark so far
            auto best = decltype(match_finish(pre))(0);
            S_t pse = 0;
            best_pse[r] = 0;
            bestj[r] = -1;
            if (cnt > 0) {
                load_staged<T>(s_rec + (unsigned)w * kRecBytes, L);
                L_j = H[r].c0;
                pre = match_prepare(L, px, py, pth, ob_beta[r], ob_r[r], ob_g[r], ob_b[r], ob_dx[r], ob_dy[r], A.prm,
                                    st_flags);
                pse = pre.pse;
            }
            st_eval += __popc(__ballot_sync(kFull, cnt > 0));
            // An item with a single colour-compatible landmark only needs the reference's decision
            // `probability > 0.0` (:369-381), and match_prepare can usually PROVE it without the two logs
            // and two exps of the pdf tails (MatchPre::sure).  The value itself is needed only to rank
            // several candidates, or when the product may underflow (finding F3) -- a warp-uniform branch.
#if PK_MATCH_SKIP == 0
            const bool need_any = true;
#else
            const bool need_any = __any_sync(kFull, cnt > 1 || (cnt == 1 && !pre.sure));
#endif
            if (need_any) {
                if (cnt > 0) {
                    const auto Lk = match_finish(pre);
                    if (Lk > 0) {
                        best = Lk;
                        bestj[r] = H[r].c0;
                        best_pse[r] = pse;
                    }
                }
            } else if (cnt > 0) {
                best = 1;  // some positive value: never compared against anything
                bestj[r] = H[r].c0;
                best_pse[r] = pse;
            }
            // further colour-compatible landmarks, in slot order.  All warp collectives sit outside the per-lane
            // conditions; lanes with more hits than the hit list holds re-walk their keys afterwards, on their own.
            const int ncand = cnt <= kMaxHits ? cnt : 0;
            const int maxcnt = __reduce_max_sync(kFull, ncand);
            // candidate record: its own registers when L must keep the best one so far (kKeepWinner), else L itself
            LM Lc_own;
            LM& Lc = kKeepWinner ? Lc_own : L;
            // candidate ring: request candidate c (>= 2) of this lane's item into strip c % 3.  Strip 0 is free (the
            // first hit is in registers), strip 1 holds the second hit, strip 2 is the ring's own.  Every call commits
            // exactly one cp.async group, so `wait_group 1` in iteration c leaves only candidate c + 1 in flight.
            auto ring_fetch = [&](int c) {
                if (c < ncand) {
                    const unsigned char* src = block + hot + (size_t)S.more_hits[par][w][c - 2] * kRecBytes;
                    const uint32_t dst = s_rec + ((unsigned)(c % 3) * 32u * R + (unsigned)w) * kRecBytes;
#pragma unroll
                    for (int i = 0; i < kChunksPerRec; ++i) cp_async16_a(dst + 16u * i, src + 16 * i);
                }
                cp_async_commit();
            };
            if (kRing && maxcnt > 2) {
                ring_fetch(2);
                ring_fetch(3);
            }
            for (int c = 1; c < maxcnt; ++c) {
                // Most extra hits are false positives of the byte-key screen: apply the exact colour
                // gate (:441) first and run the full likelihood only if some lane still needs it.
                bool need = false;
                int j = -1;
                if (kRing && c >= 2) cp_async_wait<1>();  // candidate c has landed (own copies: no warp barrier needed)
                if (c < ncand) {
                    j = (c == 1) ? H[r].c1 : (int)S.more_hits[par][w][c - 2];
                    if (c < kStaged)
                        load_staged<T>(s_rec + ((unsigned)c * 32u * R + (unsigned)w) * kRecBytes, Lc);
                    else if (kRing)
                        load_staged<T>(s_rec + ((unsigned)(c % 3) * 32u * R + (unsigned)w) * kRecBytes, Lc);
                    else
                        load_landmark<T>(block, cap, j, Lc);
                    if (!kKeepWinner) L_j = j;
                    const S_t dr = ob_r[r] - Lc.r, dg = ob_g[r] - Lc.g, db = ob_b[r] - Lc.b;
                    need = !(fabs((double)(dr * dr + dg * dg + db * db)) > A.prm.color_gate);
                }
                if (kRing && c >= 2) ring_fetch(c + 2);  // into the strip candidate c - 1 has left
                if (!__any_sync(kFull, need)) continue;
                st_eval += __popc(__ballot_sync(kFull, need));
                if (need) {
                    const auto Lk = match_likelihood(Lc, px, py, pth, ob_beta[r], ob_r[r], ob_g[r], ob_b[r],
                                                     ob_dx[r], ob_dy[r], A.prm, st_flags, pse);
                    if (Lk > best) {
                        best = Lk;
                        bestj[r] = j;
                        best_pse[r] = pse;
                        if (kKeepWinner) {
                            L = Lc;
                            L_j = j;
                        }
                    }
                }
            }
            unsigned extra_evals = 0;
            if (cnt > kMaxHits) {
                // More colour-compatible landmarks than hit registers: walk the particle's keys again
                // (from global memory this time) and evaluate every landmark that passes the key screen
                // and the exact colour gate, in slot order.
                best = 0;
                bestj[r] = -1;
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
                        load_landmark<T>(block, cap, j, Lc);
                        if (!kKeepWinner) L_j = j;
                        const auto Lk = match_likelihood(Lc, px, py, pth, ob_beta[r], ob_r[r], ob_g[r], ob_b[r],
                                                         ob_dx[r], ob_dy[r], A.prm, st_flags, pse);
                        extra_evals += 1;
                        if (Lk > best) {
                            best = Lk;
                            bestj[r] = j;
                            best_pse[r] = pse;
                            if (kKeepWinner) {
                                L = Lc;
                                L_j = j;
                            }
                        }
                    }
                }
            }
            if (__any_sync(kFull, cnt > kMaxHits)) st_eval += __reduce_add_sync(kFull, extra_evals);
        }
        // the staged records are consumed: request the next group's into the same staging area
        if (have_next) prefetch(it + 1, Hn);
        // ---- sequential updates (:88-124) in scan order -------------------------------------------
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int w = r * 32 + lane;
            const bool act = w < nitems;
            const int pl = act ? it_pl[r] : 0;
            unsigned char* block = A.pool + blk_off(S.slot_s[gi][pl]);
            const double px = S.pose[gi][pl][0], py = S.pose[gi][pl][1], pth = S.pose[gi][pl][2];
            const bool matched = act && bestj[r] >= 0;
            // items on the same landmark of the same particle go one after the other (finding F2); that can
            // only happen when two blobs of the frame have close colours
            int rank = 0, maxrank = 0;
            if (twins) {
                const int key = matched ? (pl * cap + bestj[r]) : (-1 - lane);
                const unsigned peers = __match_any_sync(kFull, key);
                rank = __popc(peers & lt);
                maxrank = __reduce_max_sync(kFull, matched ? rank : 0);
            }
            double factor = log_w ? log_no_match : A.prm.no_match_weight;  // :95 / :851-857
            int id_out = 0;
            int promoted = 0;
            for (int q = 0;; ++q) {
                if (matched && rank == q) {
                    // The winner's PRE-update record is normally still in registers; after an earlier blob of this
                    // frame rewrote the landmark (q > 0; with two rounds of items, a blob of the first round may have
                    // done so as well) it is read back from global memory.
                    const bool stale = q > 0 || (R > 1 && r > 0 && twins);
                    if (stale || R > 1 || L_j != bestj[r]) load_landmark<T>(block, cap, bestj[r], L);
                    bool changed = false;
                    // L holds the stored colours here, so its key is the one in the hot region: the fp32 instantiation
                    // skips the key store when the update leaves the key alone (a 4-byte store dirties a 32-byte sector)
                    const unsigned key_before = (sizeof(LM) == sizeof(LandmarkF)) ? stored_key(L, Rec<T>::kDtype) : kNoKey;
                    // the bearing computed during association is that of the PRE-update landmark: it is re-used
                    // unless an earlier blob of this frame has moved the landmark since
                    factor = ekf_update_lm(L, px, py, ob_beta[r], ob_r[r], ob_g[r], ob_b[r], A.prm, id_out, st_flags,
                                           promoted, changed, !stale, best_pse[r], pth, true, log_no_match);
                    if (changed) store_landmark<T>(block, cap, bestj[r], L, key_before);
                }
                if (q >= maxrank) break;
                __syncwarp();
            }
            if (maxrank > 0) st_same += __popc(__ballot_sync(kFull, matched && rank > 0));
            if (act) (A.assoc + (size_t)p0 * K)[w] = id_out;
            {
                const unsigned bm = __ballot_sync(kFull, matched), ba = __ballot_sync(kFull, act);
                st_matched += __popc(bm);
                st_unmatched += __popc(ba & ~bm);
                const unsigned bp = __ballot_sync(kFull, promoted != 0);
                if (bp) st_promoted += __popc(bp);
            }
            if (act) S.factor[w] = factor;
            // add_orphaned_reading bumps next_id once per unseen blob (:745-746)
            const unsigned unseen = __ballot_sync(kFull, act && id_out == 0);
            if (R == 1) {
                __syncwarp();
                // particles[i].weight = 1 (:73); weight *= factor in scan order (:95, :124): the first lane of a
                // particle folds its K factors left to right
                if (act && it_k[0] == 0) {
                    const int base = pl * K;
                    double wgt = log_w ? 0.0 : 1.0;
                    if (log_w) {  // PK_MODEL_LOG_WEIGHTS: the factors are logarithms, the weight is their sum
                        for (int k2 = 0; k2 < K; ++k2) wgt += S.factor[base + k2];
                    } else if (K == 8) {  // the usual scan size: four 16-byte loads, eight multiplications
                        const double2* f2 = reinterpret_cast<const double2*>(&S.factor[base]);
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) {
                            const double2 v = f2[k2];
                            wgt *= v.x;
                            wgt *= v.y;
                        }
                    } else {
                        for (int k2 = 0; k2 < K; ++k2) wgt *= S.factor[base + k2];
                    }
                    if (log_w ? (wgt != wgt || wgt == INFINITY) : !isfinite(wgt)) st_flags |= PK_FLAG_NONFINITE_WEIGHT;
                    A.pose4[4 * (size_t)(p0 + pl) + 3] = wgt;
                    const int orphans = __popc((unseen >> base) & (K >= 32 ? 0xffffffffu : ((1u << K) - 1u)));
                    if (orphans) A.aux2[2 * (size_t)(p0 + pl) + 1] += orphans;
                }
            } else {
                // one particle per group: count the unseen blobs of both rounds in lane 0
                if (lane == 0) S.unseen_acc = (r == 0 ? 0 : S.unseen_acc) + __popc(unseen);
                __syncwarp();  // also orders this round's record stores before the next round's loads
            }
        }
        if (R > 1) {
            __syncwarp();
            if (lane == 0 && gpn > 0) {
                double wgt = log_w ? 0.0 : 1.0;
                for (int k = 0; k < K; ++k) wgt = log_w ? wgt + S.factor[k] : wgt * S.factor[k];
                if (log_w ? (wgt != wgt || wgt == INFINITY) : !isfinite(wgt)) st_flags |= PK_FLAG_NONFINITE_WEIGHT;
                A.pose4[4 * (size_t)p0 + 3] = wgt;
                const int orphans = S.unseen_acc;
                if (orphans) A.aux2[2 * (size_t)p0 + 1] += orphans;
            }
        }
        __syncwarp();
    };

    // ---- software pipeline over this warp's groups -----------------------------------------------
    // scan(it + 1) -> association(it) -> record prefetch(it + 1) -> updates(it): keys run one group ahead through
    // the key ring, records one phase ahead through the (single) staging area.
    Hits Hcur[R], Hnext[R];
#pragma unroll
    for (int r = 0; r < R; ++r) Hcur[r] = Hnext[r] = Hits{0, -1, -1};
    for (int s = 0; s < kStages - 1; ++s) produce_one();
    cp_async_commit();  // stands for the records of a previous group: scan() waits for all but the newest group
    if (my_groups > 0) {
        scan(0, Hcur);
        prefetch(0, Hcur);
    }
    for (int it = 0; it < my_groups; ++it) {
        const bool have_next = it + 1 < my_groups;
        if (have_next) scan(it + 1, Hnext); else cp_async_commit();  // (keeps the group count evaluate() waits on)
        evaluate(it, Hcur, Hnext, have_next);
#pragma unroll
        for (int r = 0; r < R; ++r) Hcur[r] = Hnext[r];
    }

    // ---- statistics: one atomic per warp per counter ------------------------------------------------
    st_flags = __reduce_or_sync(kFull, st_flags);
    if (lane == 0 && A.stats != nullptr) {
        if (st_matched) atomicAdd(&A.stats[PK_STAT_MATCHED], (unsigned long long)st_matched);
        if (st_unmatched) atomicAdd(&A.stats[PK_STAT_UNMATCHED], (unsigned long long)st_unmatched);
        if (st_eval) atomicAdd(&A.stats[PK_STAT_EVALUATED], (unsigned long long)st_eval);
        if (st_same) atomicAdd(&A.stats[PK_STAT_SAME_LANDMARK], (unsigned long long)st_same);
        if (st_promoted) atomicAdd(&A.stats[PK_STAT_PROMOTED], (unsigned long long)st_promoted);
        if (st_flags) atomicOr(&A.stats[PK_STAT_FLAGS], (unsigned long long)st_flags);
    }
}

__global__ void reset_weight_kernel(double* __restrict__ pose4, long long M, double value) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) pose4[4 * i + 3] = value;  // cam_cb :73 with an empty scan (log mode: log 1)
}

template <typename T, int R, typename LM>
static int launch_measure(MeasureArgs& args, cudaStream_t st) {
    auto align128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    args.keys_off = (int)align128(sizeof(WarpSmemT<R>));
    args.rec_off = (int)align128(args.keys_off + (size_t)kStages * args.group * kKeyStride * 4);
    args.warp_smem = (int)align128(args.rec_off + (size_t)staging_strips<T, LM>() * 32 * R * sizeof(typename Rec<T>::Cold));
    const size_t smem = (size_t)args.warp_smem * kWarpsPerCta;
    // the attribute is per device (a process may run filters on several): set it on every launch, it is cheap
    if (smem > 48 * 1024)
        PK_CUDA(cudaFuncSetAttribute(measure_kernel<T, R, LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = 0;
    PK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, measure_kernel<T, R, LM>, kWarpsPerCta * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const long long n_groups = ((long long)args.M + args.group - 1) / args.group;
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
__global__ void obs_table_kernel(const double* __restrict__ obs, int K, double color_gate, pk::ObsTable* __restrict__ tab) {
    const int k = threadIdx.x;
    if (k >= PK_MAX_OBS) return;
    // does any other blob have a colour close enough to share a landmark with blob k?  (ObsTable::twins)
    bool tw = false;
    if (k < K)
        for (int k2 = 0; k2 < K; ++k2)
            if (k2 != k && pk::blobs_may_share_landmark(obs[4 * k + 1], obs[4 * k + 2], obs[4 * k + 3], obs[4 * k2 + 1],
                                                        obs[4 * k2 + 2], obs[4 * k2 + 3], color_gate))
                tw = true;
    const int any_tw = __syncthreads_or(tw ? 1 : 0);
    if (k == 0) tab->twins = any_tw ? 1u : 0u;
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
    PK_CHECK_ARG(M >= 0 && M < (1ll << 31) - 64, "M must be in [0, 2^31 - 64)");
    PK_CHECK_ARG(K >= 0 && K <= PK_MAX_OBS, "K must be in [0, PK_MAX_OBS]");
    PK_CHECK_ARG(capacity >= 0 && capacity < (1 << 16), "capacity must be < 2^16");
    // record positions travel between lanes as 32-bit offsets in units of 32 bytes
    PK_CHECK_ARG((double)M * (double)block_bytes(capacity, dtype) < 137438953472.0, "landmark pool must be < 2^37 bytes");
    PK_CHECK_ARG(params != nullptr, "params is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) return PK_OK;
    if (K == 0) {
        const int threads = 256;
        reset_weight_kernel<<<(unsigned)((M + threads - 1) / threads), threads, 0, st>>>(
            pose4, M, (params->model & PK_MODEL_LOG_WEIGHTS) ? 0.0 : 1.0);
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
    args.M = (int)M;
    args.block_bytes = (unsigned)block_bytes(capacity, dtype);
    args.hot_bytes = (unsigned)hot_region_bytes(capacity);
    args.capacity = capacity;
    args.K = K;
    int group = 32 / K;
    if (group < 1) group = 1;
    if (group > kMaxGroup) group = kMaxGroup;
    args.group = group;
    args.prm = *params;
    args.log_no_match = log(params->no_match_weight);
    args.tab_dev = nullptr;
    if (obs_dev != nullptr) {
        PK_CHECK_ARG(table_ws != nullptr, "table workspace is NULL");
        obs_table_kernel<<<1, PK_MAX_OBS, 0, st>>>(obs_dev, K, params->color_gate, (ObsTable*)table_ws);
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
        T.twins = 0u;
        for (int k = 0; k < K; ++k)
            for (int k2 = k + 1; k2 < K; ++k2)
                if (blobs_may_share_landmark(T.cr[k], T.cg[k], T.cb[k], T.cr[k2], T.cg[k2], T.cb[k2], params->color_gate))
                    T.twins = 1u;
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
    // sum |d| <= sqrt(3 * sum d^2): the largest integer t with t^2 <= 3 * key_thr (-1: nothing passes)
    {
        const long long three = 3ll * (long long)args.key_thr;
        long long t = three < 0 ? -1 : (long long)floor(sqrt((double)three));
        while (three >= 0 && (t + 1) * (t + 1) <= three) ++t;
        while (t >= 0 && t * t > three) --t;
        args.key_thr1 = (int)t;
    }
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
