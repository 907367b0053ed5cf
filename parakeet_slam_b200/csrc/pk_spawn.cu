// K2b -- spawn mode: new-landmark hypotheses for the unseen blobs of a frame.
//
// Replaces FilterParticle.add_hypothesis (reference prkt_core_v2.py:546-563) and what it calls:
// find_nearest_reading (:565-590), reading_distance_function (:592-608), ray_intersect (:610-640),
// color_distance (:642-651), add_new_feature (:653-680), cross_readings (:682-738) and
// add_orphaned_reading (:740-746).  As written that path is dead code (SURVEY.md finding F5: the
// search iterates the wrong dict); this kernel implements the reference with the three patches of
// SURVEY.md A.6 (oracle/ref_shim.apply_spawn_patches):
//   P1  find_nearest_reading iterates hypothesis_set in insertion order, strict '<' (first minimum
//       wins), and pairs when the minimum colour distance is <= pair_gate;
//   P2  an orphaned reading stores a COPY of the pose;
//   P3  the intersection is two plain floats.
// Reported deviation: readings live in a ring of `slots` entries per particle (the reference never
// forgets one); an overwrite sets PK_FLAG_ORPHAN_EXPIRED.
//
// Shape: unseen blobs are rare once the map is known, so a warp first looks at 32 particles at once
// (one lane each: the particle's row of association ids) and then serves, one after the other, only
// those that have an unseen blob.  For such a particle the lanes are the stored readings: every lane
// tests its reading's ray against the new one and the warp takes the arg-min.  The arithmetic uses
// explicit round-to-nearest intrinsics (no FMA contraction) so it follows the reference operation
// by operation; cos / sin are CUDA's (<= 2 ulp from libm's).
#include <math.h>

#include "pk_common.cuh"

namespace pk {

constexpr unsigned kFullMaskS = 0xffffffffu;

struct SpawnArgs {
    const double* pose4;
    int* aux2;
    const int* slot;
    unsigned char* pool;
    const int* assoc;
    unsigned long long* stats;
    long long M;
    size_t block_bytes;
    size_t orph_off;
    int capacity;
    int K;
    int slots;
    double gate;
    const double* obs_dev;  // device-resident scan [K][4] (pk_spawn_update_dev), else NULL
    double beta[PK_MAX_OBS], cr[PK_MAX_OBS], cg[PK_MAX_OBS], cb[PK_MAX_OBS];
};

struct Reading {
    double x, y, c, s, r, g, b, id;
};

__device__ __forceinline__ Reading load_reading(const unsigned char* p) {
    const int4 a = ldcg16(p), b = ldcg16(p + 16), c = ldcg16(p + 32), d = ldcg16(p + 48);
    return Reading{i4lo(a), i4hi(a), i4lo(b), i4hi(b), i4lo(c), i4hi(c), i4lo(d), i4hi(d)};
}
__device__ __forceinline__ void store_reading(unsigned char* p, const Reading& q) {
    stcg16(p, mk_i4(q.x, q.y));
    stcg16(p + 16, mk_i4(q.c, q.s));
    stcg16(p + 32, mk_i4(q.r, q.g));
    stcg16(p + 48, mk_i4(q.b, q.id));
}

// ray_intersect :610-640 with a = the stored reading (state1, blob1), b = the new one
__device__ __forceinline__ bool ray_intersect(double as0, double as1, double ad0, double ad1, double bs0, double bs1,
                                              double bd0, double bd1) {
    const double den = __dsub_rn(__dmul_rn(ad1, bd0), __dmul_rn(ad0, bd1));
    if (den == 0.0) return false;
    // v = ((ad0*bs1 - ad1*bs0 + ad1*as0 - ad0*as1) / den), evaluated left to right
    const double num = __dsub_rn(__dadd_rn(__dsub_rn(__dmul_rn(ad0, bs1), __dmul_rn(ad1, bs0)), __dmul_rn(ad1, as0)),
                                 __dmul_rn(ad0, as1));
    const double v = __ddiv_rn(num, den);
    double u;
    if (fabs(ad1) < fabs(ad0))
        u = __ddiv_rn(__dsub_rn(__dadd_rn(bs0, __dmul_rn(bd0, v)), as0), ad0);
    else
        u = __ddiv_rn(__dsub_rn(__dadd_rn(bs1, __dmul_rn(bd1, v)), as1), ad1);
    return u >= 0.0 && v >= 0.0;
}

// cross_readings :682-738 (line-line intersection through two points per ray); false if "None"
__device__ __forceinline__ bool cross_readings(double x1, double y1, double c1, double s1, double x3, double y3, double c3,
                                               double s3, double& X, double& Y) {
    const double x2 = __dadd_rn(x1, c1), y2 = __dadd_rn(y1, s1);
    const double x4 = __dadd_rn(x3, c3), y4 = __dadd_rn(y3, s3);
    const double t0 = __dsub_rn(__dmul_rn(x1, y2), __dmul_rn(y1, x2));
    const double t1 = __dsub_rn(x3, x4);
    const double t2 = __dsub_rn(x1, x2);
    const double t3 = __dsub_rn(__dmul_rn(x3, y4), __dmul_rn(x4, y3));
    const double t5 = __dsub_rn(y3, y4);
    const double t6 = __dsub_rn(y1, y2);
    const double den = __dsub_rn(__dmul_rn(t2, t5), __dmul_rn(t6, t1));
    if (den == 0.0) return false;
    X = __ddiv_rn(__dsub_rn(__dmul_rn(t0, t1), __dmul_rn(t2, t3)), den);
    Y = __ddiv_rn(__dsub_rn(__dmul_rn(t0, t5), __dmul_rn(t6, t3)), den);
    return true;
}

template <typename T>
__global__ void __launch_bounds__(128)
spawn_kernel(const __grid_constant__ SpawnArgs A) {
    const int lane = threadIdx.x & 31;
    const long long total_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int K = A.K, slots = A.slots, cap = A.capacity;
    unsigned long long st_spawned = 0, st_orphaned = 0;
    unsigned st_flags = 0;

    for (long long base = gw * 32; base < A.M; base += total_warps * 32) {
        const long long p = base + lane;
        unsigned long long mask = 0ull;  // bit k: blob k of my particle is unseen (id 0, :91)
        if (p < A.M)
            for (int k = 0; k < K; ++k)
                if (A.assoc[p * K + k] == 0) mask |= 1ull << k;
        unsigned need = __ballot_sync(kFullMaskS, mask != 0ull);
        while (need) {
            const int src = __ffs((int)need) - 1;
            need &= need - 1u;
            const long long q = base + src;
            const unsigned long long qmask = __shfl_sync(kFullMaskS, mask, src);
            unsigned char* block = A.pool + (size_t)A.slot[q] * A.block_bytes;
            unsigned char* oh = block + A.orph_off;
            unsigned char* recs = oh + kOrphanHeaderBytes;
            int total = __ldcg(reinterpret_cast<const int*>(oh));
            const double px = A.pose4[4 * q], py = A.pose4[4 * q + 1], pth = A.pose4[4 * q + 2];
            int n_live = A.aux2[2 * q];
            // pk_measurement_update already counted every unseen blob into next_id (:745-746 / :680)
            int id_next = A.aux2[2 * q + 1] - __popcll(qmask);
            for (unsigned long long rest = qmask; rest; rest &= rest - 1ull) {
                const int k = __ffsll((long long)rest) - 1;
                const double beta = A.obs_dev ? A.obs_dev[4 * k] : A.beta[k];
                const double orr = A.obs_dev ? A.obs_dev[4 * k + 1] : A.cr[k];
                const double og = A.obs_dev ? A.obs_dev[4 * k + 2] : A.cg[k];
                const double ob = A.obs_dev ? A.obs_dev[4 * k + 3] : A.cb[k];
                // world-frame ray of the new reading: b2 = blob2.bearing + heading (:603)
                const double ang = __dadd_rn(beta, pth);
                double bd0, bd1;
                sincos(ang, &bd1, &bd0);
                const int live = min(total, slots);
                const int start = (total > slots) ? (total % slots) : 0;
                // find_nearest_reading (P1): insertion order, strict '<'
                double best_d = INFINITY;
                int best_i = -1;
                for (int i = lane; i < live; i += 32) {
                    const Reading R = load_reading(recs + (size_t)((start + i) % slots) * kOrphanBytes);
                    double d = INFINITY;  // reading_distance_function :592-608
                    if (ray_intersect(R.x, R.y, R.c, R.s, px, py, bd0, bd1)) {
                        const double dr = __dsub_rn(R.r, orr), dg = __dsub_rn(R.g, og), db = __dsub_rn(R.b, ob);
                        d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dg, dg)), __dmul_rn(db, db)));  // :642-651
                    }
                    if (d < best_d) {
                        best_d = d;
                        best_i = i;
                    }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double od = __shfl_xor_sync(kFullMaskS, best_d, o);
                    const int oi = __shfl_xor_sync(kFullMaskS, best_i, o);
                    if (oi >= 0 && (best_i < 0 || od < best_d || (od == best_d && oi < best_i))) {
                        best_d = od;
                        best_i = oi;
                    }
                }
                bool pair = best_i >= 0 && best_d <= A.gate;  // pair_id > 0 (:557-560)
                if (lane == 0) {
                    if (pair) {
                        // add_new_feature :653-680
                        const Reading R = load_reading(recs + (size_t)((start + best_i) % slots) * kOrphanBytes);
                        double X = 0.0, Y = 0.0;
                        if (!cross_readings(R.x, R.y, R.c, R.s, px, py, bd0, bd1, X, Y)) {
                            st_flags |= PK_FLAG_SPAWN_DEGENERATE;
                            pair = false;
                        } else if (n_live >= cap) {
                            st_flags |= PK_FLAG_MAP_FULL;  // dropped; the id is consumed all the same
                        } else {
                            Landmark L;
                            L.x = X;
                            L.y = Y;
                            L.r = __ddiv_rn(__dadd_rn(R.r, orr), 2.0);
                            L.g = __ddiv_rn(__dadd_rn(R.g, og), 2.0);
                            L.b = __ddiv_rn(__dadd_rn(R.b, ob), 2.0);
                            L.sp[0] = 1.0; L.sp[1] = 0.0; L.sp[2] = 0.0; L.sp[3] = 1.0;
#pragma unroll
                            for (int e = 0; e < 9; ++e) L.sc[e] = (e % 4 == 0) ? 1.0 : 0.0;
                            L.meta = PK_META_POTENTIAL;  // update_count 0, lives in potential_features (:679)
                            L.id = -id_next;
                            store_landmark<T>(block, cap, n_live, L);
                            n_live += 1;
                            st_spawned += 1;
                        }
                    }
                    if (!pair) {
                        // add_orphaned_reading :740-746 (P2: the pose is copied)
                        Reading N{px, py, bd0, bd1, orr, og, ob, (double)id_next};
                        store_reading(recs + (size_t)(total % slots) * kOrphanBytes, N);
                        if (total >= slots) st_flags |= PK_FLAG_ORPHAN_EXPIRED;
                        st_orphaned += 1;
                    }
                }
                pair = __shfl_sync(kFullMaskS, (int)pair, 0) != 0;
                if (!pair) total += 1;
                id_next += 1;  // :680 / :746
                __syncwarp();
            }
            if (lane == 0) {
                __stcg(reinterpret_cast<int*>(oh), total);
                A.aux2[2 * q] = n_live;
            }
            __syncwarp();
        }
    }
    for (int o = 16; o > 0; o >>= 1) st_flags |= __shfl_xor_sync(kFullMaskS, st_flags, o);
    if (lane == 0 && A.stats != nullptr) {
        if (st_spawned) atomicAdd(&A.stats[PK_STAT_SPAWNED], st_spawned);
        if (st_orphaned) atomicAdd(&A.stats[PK_STAT_ORPHANED], st_orphaned);
        if (st_flags) atomicOr(&A.stats[PK_STAT_FLAGS], (unsigned long long)st_flags);
    }
}

__global__ void orphans_export_kernel(const unsigned char* __restrict__ pool, size_t bbytes, size_t orph_off, int slots,
                                      const int* __restrict__ slot, long long p_lo, long long count,
                                      int* __restrict__ totals, double* __restrict__ readings) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * slots) return;
    const long long pi = t / slots;
    const int j = (int)(t % slots);
    const unsigned char* oh = pool + (size_t)slot[p_lo + pi] * bbytes + orph_off;
    if (j == 0) totals[pi] = *reinterpret_cast<const int*>(oh);
    const double* src = reinterpret_cast<const double*>(oh + kOrphanHeaderBytes + (size_t)j * kOrphanBytes);
    for (int e = 0; e < 8; ++e) readings[t * 8 + e] = src[e];
}

}  // namespace pk

using namespace pk;

extern "C" {

static int spawn_common(const double* pose4, int* aux2, const int* slot, void* pool, int capacity, int dtype, long long M,
                        const double* obs_host, const double* obs_dev, int K, const int* assoc, double pair_gate,
                        unsigned long long* stats, void* stream) {
    PK_CHECK_ARG(pose4 && aux2 && slot && pool, "null state pointer");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    PK_CHECK_ARG(dtype_orphans(dtype) > 0, "dtype carries no orphan slots (PK_DTYPE_WITH_ORPHANS)");
    PK_CHECK_ARG(M >= 0, "M < 0");
    PK_CHECK_ARG(K >= 0 && K <= PK_MAX_OBS, "K must be in [0, PK_MAX_OBS]");
    PK_CHECK_ARG(capacity >= 0, "capacity");
    if (M == 0 || K == 0) return PK_OK;
    PK_CHECK_ARG((obs_host != nullptr || obs_dev != nullptr) && assoc != nullptr, "obs / assoc is NULL");
    static thread_local SpawnArgs args;
    args.obs_dev = obs_dev;
    args.pose4 = pose4;
    args.aux2 = aux2;
    args.slot = slot;
    args.pool = (unsigned char*)pool;
    args.assoc = assoc;
    args.stats = stats;
    args.M = M;
    args.block_bytes = block_bytes(capacity, dtype);
    args.orph_off = orphan_offset(capacity, dtype);
    args.capacity = capacity;
    args.K = K;
    args.slots = dtype_orphans(dtype);
    args.gate = pair_gate;
    for (int k = 0; k < K && obs_host != nullptr; ++k) {
        args.beta[k] = obs_host[4 * k + 0];
        args.cr[k] = obs_host[4 * k + 1];
        args.cg[k] = obs_host[4 * k + 2];
        args.cb[k] = obs_host[4 * k + 3];
    }
    const int threads = 128;
    long long grid = (M + 32ll * (threads / 32) - 1) / (32ll * (threads / 32));
    const long long cap_grid = (long long)num_sms() * 16;
    if (grid > cap_grid) grid = cap_grid;
    if (grid < 1) grid = 1;
    if (dtype_base(dtype) == PK_DTYPE_F32)
        spawn_kernel<float><<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(args);
    else
        spawn_kernel<double><<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(args);
    PK_LAUNCH_CHECK("spawn_kernel");
    return PK_OK;
}

int pk_spawn_update(const double* pose4, int* aux2, const int* slot, void* pool, int capacity, int dtype, long long M,
                    const double* obs_host, int K, const int* assoc, double pair_gate, unsigned long long* stats,
                    void* stream) {
    PK_CHECK_ARG(K == 0 || M == 0 || obs_host != nullptr, "obs_host is NULL");
    return spawn_common(pose4, aux2, slot, pool, capacity, dtype, M, obs_host, nullptr, K, assoc, pair_gate, stats, stream);
}

int pk_spawn_update_dev(const double* pose4, int* aux2, const int* slot, void* pool, int capacity, int dtype, long long M,
                        const double* obs_dev, int K, const int* assoc, double pair_gate, unsigned long long* stats,
                        void* stream) {
    PK_CHECK_ARG(K == 0 || M == 0 || obs_dev != nullptr, "obs_dev is NULL");
    return spawn_common(pose4, aux2, slot, pool, capacity, dtype, M, nullptr, obs_dev, K, assoc, pair_gate, stats, stream);
}

int pk_orphans_export(const void* pool, int capacity, int dtype, const int* slot, long long p_lo, long long count,
                      int* totals, double* readings, void* stream) {
    PK_CHECK_ARG(pool && slot && totals && readings, "null pointer");
    PK_CHECK_ARG(dtype_valid(dtype) && dtype_orphans(dtype) > 0, "dtype carries no orphan slots");
    PK_CHECK_ARG(count >= 0 && p_lo >= 0, "sizes");
    if (count == 0) return PK_OK;
    const int slots = dtype_orphans(dtype);
    const long long total = count * slots;
    orphans_export_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)pool, block_bytes(capacity, dtype), orphan_offset(capacity, dtype), slots, slot, p_lo, count,
        totals, readings);
    PK_LAUNCH_CHECK("orphans_export_kernel");
    return PK_OK;
}

}  // extern "C"
