// Shared device/host helpers for libparakeet_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/parakeet_b200.h"

#ifndef __CUDA_ARCH__
#include <stdio.h>
#include <string.h>
#endif

namespace pk {

// ---------------------------------------------------------------------------------------------
// Error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define PK_CHECK_ARG(cond, msg)                         \
    do {                                                \
        if (!(cond)) {                                  \
            pk::set_error("invalid argument: %s", msg); \
            return PK_EINVAL;                           \
        }                                               \
    } while (0)

#define PK_CUDA(call)                                         \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return pk::cuda_fail(_e, #call); \
    } while (0)

#define PK_LAUNCH_CHECK(name)                                      \
    do {                                                           \
        cudaError_t _e = cudaGetLastError();                       \
        if (_e != cudaSuccess) return pk::cuda_fail(_e, name);     \
    } while (0)

int num_sms();

// ---------------------------------------------------------------------------------------------
// Landmark record layout.  One particle's map ("block") = [hot region][cold region].
//
//   hot   4 B / landmark: a colour KEY -- r,g,b rounded and clamped to bytes (top byte zero).
//         It is all the association pre-filter needs and the only part of the map the fused
//         kernel streams for every landmark (256 B per particle at 64 landmarks).
//   cold  64 B (f32) / 160 B (f64) per landmark: colour mean, position mean, the two covariance
//         blocks, id and update_count | flags.  Fetched only for the few landmarks whose key
//         survives the colour gate.  The f32 record is exactly one 64-byte DRAM granule: it keeps
//         the covariance blocks as their LOWER triangles (the triangle the reference's pdf reads,
//         scipy eigh(lower=True)); the two off-diagonal copies differ only by rounding in the
//         reference, far below fp32 resolution.  The f64 record keeps all 13 entries.
//
// 4 + 64 = 68 B per landmark in f32 and 4 + 160 = 164 B in f64 (SURVEY.md 8(d) budgeted 84 / 164).
// ---------------------------------------------------------------------------------------------
struct alignas(16) ColdF {
    float r, g, b, x, y;
    float sp[3];  // position covariance block, lower triangle: S00, S10, S11
    float sc[6];  // colour covariance block, lower triangle: C00, C10, C11, C20, C21, C22
    int id;       // reference landmark id (>0 full, <0 potential)
    int meta;     // Feature.update_count | PK_META_IMMUTABLE | PK_META_POTENTIAL
};
struct alignas(16) ColdD {
    double r, g, b, x, y;
    double sp[4];
    double sc[9];
    int id;
    int meta;
    int pad[2];
};
static_assert(sizeof(ColdF) == 64 && sizeof(ColdD) == 160, "cold record layout");

template <typename T>
struct Rec;
template <>
struct Rec<float> {
    using Cold = ColdF;
    static constexpr int kDtype = PK_DTYPE_F32;
};
template <>
struct Rec<double> {
    using Cold = ColdD;
    static constexpr int kDtype = PK_DTYPE_F64;
};

// Per-frame blob table of the measurement kernels: built on the host from obs_host (kernel argument, no
// copy) or by obs_table_kernel from a device-resident scan (pk_measurement_update_dev).
struct ObsTable {
    double beta[PK_MAX_OBS], cr[PK_MAX_OBS], cg[PK_MAX_OBS], cb[PK_MAX_OBS];
    double dirx[PK_MAX_OBS], diry[PK_MAX_OBS];  // unit((cos b, sin b, 0)) of closest_point :510
    unsigned okey[PK_MAX_OBS];                  // colour keys of the blobs
    // != 0 when some pair of blobs has colours within twice the gate radius of each other: only then can two
    // blobs of one frame pass the colour gate (:441) of the SAME landmark (triangle inequality), i.e. only then
    // does the fused kernel have to order updates of one landmark (finding F2)
    unsigned twins;
};
// squared colour distance below which two blobs count as twins (conservative: rounding of the fp32 gate, NaN gate)
__host__ __device__ inline bool blobs_may_share_landmark(double r1, double g1, double b1, double r2, double g2, double b2,
                                                          double gate) {
    if (!(gate == gate)) return true;  // NaN gate: everything passes
    if (gate < 0.0) return false;      // nothing passes
    const double d2 = (r1 - r2) * (r1 - r2) + (g1 - g2) * (g1 - g2) + (b1 - b2) * (b1 - b2);
    return !(d2 > 4.0 * gate * 1.001 + 1.0);
}

// fp64 working copy of one landmark
struct Landmark {
    double x, y, r, g, b;
    double sp[4];
    double sc[9];
    int meta;  // update_count | PK_META_IMMUTABLE | PK_META_POTENTIAL
    int id;
};

// `dtype` arguments carry a LAYOUT CODE: the storage type in the low byte and, in spawn mode, the number
// of orphan-reading slots per particle above it (PK_DTYPE_WITH_ORPHANS).  The orphan region
// [64-byte header | n x 64-byte readings] follows the cold region inside the particle's block, so
// it is copied (and migrates between ranks) with the map.
// fp32 working copy (PK_DTYPE_ARITH_F32: fp32 landmark algebra on fp32 storage; poses, weights and the
// resampling stay fp64).  Covariance blocks keep the stored lower triangles only.
struct LandmarkF {
    float x, y, r, g, b;
    float sp[3];  // S00, S10, S11
    float sc[6];  // C00, C10, C11, C20, C21, C22
    int meta;
    int id;
};

__host__ __device__ inline int dtype_base(int dtype) { return dtype & 0xff; }
__host__ __device__ inline int dtype_orphans(int dtype) { return (dtype >> 8) & 0xffff; }
__host__ __device__ inline bool dtype_arith_f32(int dtype) { return (dtype & PK_DTYPE_ARITH_F32) != 0; }
__host__ __device__ inline bool dtype_valid(int dtype) {
    return (dtype_base(dtype) == PK_DTYPE_F32 || dtype_base(dtype) == PK_DTYPE_F64) &&
           (dtype & ~(0xffffff | PK_DTYPE_ARITH_F32)) == 0 && dtype_orphans(dtype) <= PK_MAX_ORPHANS &&
           !(dtype_arith_f32(dtype) && dtype_base(dtype) != PK_DTYPE_F32);
}
constexpr int kOrphanHeaderBytes = 64;  // int total (readings ever stored; ring position = total % slots)
constexpr int kOrphanBytes = 64;        // double x, y, cos(ray), sin(ray), r, g, b, id

__host__ __device__ inline size_t hot_bytes(int) { return 4; }
__host__ __device__ inline size_t cold_bytes(int dtype) { return dtype_base(dtype) == PK_DTYPE_F64 ? sizeof(ColdD) : sizeof(ColdF); }
// hot region padded to 64 B: every f32 cold record is then exactly one 64-byte DRAM granule and every f64 record
// starts on a 32-byte sector, so records are written with full-sector (256-bit) stores -- a 16-byte store is a
// partial-sector write whose miss in L2 costs a fill read from DRAM
__host__ __device__ inline size_t hot_region_bytes(int capacity) { return ((size_t)capacity * 4 + 63) & ~(size_t)63; }
__host__ __device__ inline size_t orphan_offset(int capacity, int dtype) {
    return hot_region_bytes(capacity) + (size_t)capacity * cold_bytes(dtype);
}
__host__ __device__ inline size_t orphan_region_bytes(int dtype) {
    const int n = dtype_orphans(dtype);
    return n ? (size_t)kOrphanHeaderBytes + (size_t)n * kOrphanBytes : 0;
}
__host__ __device__ inline size_t block_bytes(int capacity, int dtype) {
    return orphan_offset(capacity, dtype) + orphan_region_bytes(dtype);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Cache-global (L2) vector loads/stores: landmark records are written and re-read by different
// lanes of a warp inside one kernel, so they must never be served from a stale L1 line.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const int4*>(p)); }
// Same, with the L2 fetch size limited to 64 bytes.  On B200 an L2 read miss fetches the whole 128-byte line from
// DRAM by default (measured with tools/dram_probe.cu: scattered 64-byte records cost 128 bytes of DRAM reads each,
// whatever cudaLimitMaxL2FetchGranularity says); the .L2::64B qualifier halves that.  Used for every scattered
// record access; streams that cover whole lines anyway keep the default.
__device__ __forceinline__ int4 ldcg16_rec(const void* p) {
    int4 v;
    asm volatile("ld.global.cg.L2::64B.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stcg16(void* p, int4 v) { __stcg(reinterpret_cast<int4*>(p), v); }

// one full 32-byte sector per request (STG.256); p must be 32-byte aligned
__device__ __forceinline__ void stcg32(void* p, int4 a, int4 b) {
    asm volatile("st.global.cg.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

__device__ __forceinline__ double i4lo(int4 v) { return __hiloint2double(v.y, v.x); }
__device__ __forceinline__ double i4hi(int4 v) { return __hiloint2double(v.w, v.z); }
__device__ __forceinline__ int4 mk_i4(double a, double b) {
    return make_int4(__double2loint(a), __double2hiint(a), __double2loint(b), __double2hiint(b));
}

// colour key: each channel clamped to [0,255] and rounded to nearest (NaN -> 0)
__device__ __forceinline__ unsigned color_key(double r, double g, double b) {
    const unsigned kr = (unsigned)__double2int_rn(fmin(fmax(r, 0.0), 255.0));
    const unsigned kg = (unsigned)__double2int_rn(fmin(fmax(g, 0.0), 255.0));
    const unsigned kb = (unsigned)__double2int_rn(fmin(fmax(b, 0.0), 255.0));
    return kr | (kg << 8) | (kb << 16);
}
template <typename T>
__device__ __forceinline__ const unsigned char* cold_ptr(const unsigned char* block, int capacity, int j) {
    return block + hot_region_bytes(capacity) + (size_t)j * sizeof(typename Rec<T>::Cold);
}

// colour mean only (first 16 B / 32 B of the cold record) -- the exact colour gate
template <typename T>
__device__ __forceinline__ void load_colour(const unsigned char* block, int capacity, int j, double& r, double& g, double& b);
template <>
__device__ __forceinline__ void load_colour<float>(const unsigned char* block, int capacity, int j, double& r, double& g, double& b) {
    const int4 c0 = ldcg16_rec(cold_ptr<float>(block, capacity, j));
    r = (double)__int_as_float(c0.x);
    g = (double)__int_as_float(c0.y);
    b = (double)__int_as_float(c0.z);
}
template <>
__device__ __forceinline__ void load_colour<double>(const unsigned char* block, int capacity, int j, double& r, double& g, double& b) {
    const unsigned char* cp = cold_ptr<double>(block, capacity, j);
    const int4 c0 = ldcg16_rec(cp), c1 = ldcg16_rec(cp + 16);
    r = i4lo(c0);
    g = i4hi(c0);
    b = i4lo(c1);
}

// cold record (as int4 words in registers) -> fp64 working copy
template <typename T>
__device__ __forceinline__ void decode_cold(const int4* c, Landmark& L);
template <>
__device__ __forceinline__ void decode_cold<float>(const int4* c, Landmark& L) {
    L.r = (double)__int_as_float(c[0].x);
    L.g = (double)__int_as_float(c[0].y);
    L.b = (double)__int_as_float(c[0].z);
    L.x = (double)__int_as_float(c[0].w);
    L.y = (double)__int_as_float(c[1].x);
    L.sp[0] = (double)__int_as_float(c[1].y);
    L.sp[2] = (double)__int_as_float(c[1].z);
    L.sp[1] = L.sp[2];
    L.sp[3] = (double)__int_as_float(c[1].w);
    L.sc[0] = (double)__int_as_float(c[2].x);
    L.sc[3] = (double)__int_as_float(c[2].y);
    L.sc[4] = (double)__int_as_float(c[2].z);
    L.sc[6] = (double)__int_as_float(c[2].w);
    L.sc[7] = (double)__int_as_float(c[3].x);
    L.sc[8] = (double)__int_as_float(c[3].y);
    L.sc[1] = L.sc[3];
    L.sc[2] = L.sc[6];
    L.sc[5] = L.sc[7];
    L.id = c[3].z;
    L.meta = c[3].w;
}
template <>
__device__ __forceinline__ void decode_cold<double>(const int4* c, Landmark& L) {
    L.r = i4lo(c[0]);
    L.g = i4hi(c[0]);
    L.b = i4lo(c[1]);
    L.x = i4hi(c[1]);
    L.y = i4lo(c[2]);
    L.sp[0] = i4hi(c[2]);
    L.sp[1] = i4lo(c[3]);
    L.sp[2] = i4hi(c[3]);
    L.sp[3] = i4lo(c[4]);
    L.sc[0] = i4hi(c[4]);
    L.sc[1] = i4lo(c[5]);
    L.sc[2] = i4hi(c[5]);
    L.sc[3] = i4lo(c[6]);
    L.sc[4] = i4hi(c[6]);
    L.sc[5] = i4lo(c[7]);
    L.sc[6] = i4hi(c[7]);
    L.sc[7] = i4lo(c[8]);
    L.sc[8] = i4hi(c[8]);
    L.id = c[9].x;
    L.meta = c[9].y;
}

template <typename T>
__device__ __forceinline__ void load_landmark(const unsigned char* block, int capacity, int j, Landmark& L) {
    constexpr int kWords = (int)(sizeof(typename Rec<T>::Cold) / 16);
    const unsigned char* cp = cold_ptr<T>(block, capacity, j);
    int4 c[kWords];
#pragma unroll
    for (int i = 0; i < kWords; ++i) c[i] = ldcg16_rec(cp + 16 * i);
    decode_cold<T>(c, L);
}

// same, from a record staged in shared memory by a TMA bulk copy (32-bit shared address)
template <typename T>
__device__ __forceinline__ void load_staged(uint32_t rec, Landmark& L);

// kNoKey: "previous key unknown, always store it".  A 4-byte key store dirties a whole 32-byte sector (one
// fill read + one write back per sector); the key of a landmark that is merely refined almost never changes, so
// callers that know the previous key pass it and the store is skipped when it is still right.
constexpr unsigned kNoKey = 0xffffffffu;

template <typename T>
__device__ __forceinline__ void store_landmark(unsigned char* block, int capacity, int j, const Landmark& L,
                                               unsigned old_key = kNoKey);

template <>
__device__ __forceinline__ void store_landmark<float>(unsigned char* block, int capacity, int j, const Landmark& L,
                                                      unsigned old_key) {
    unsigned char* cp = const_cast<unsigned char*>(cold_ptr<float>(block, capacity, j));
#define PK_F(v) __float_as_int((float)(v))
    const float fr = (float)L.r, fg = (float)L.g, fb = (float)L.b;
    // the key is derived from the STORED (rounded) colour so screen and exact test see one value
    const unsigned key = color_key((double)fr, (double)fg, (double)fb);
    if (key != old_key) __stcg(reinterpret_cast<unsigned*>(block) + j, key);
    stcg32(cp, make_int4(__float_as_int(fr), __float_as_int(fg), __float_as_int(fb), PK_F(L.x)),
           make_int4(PK_F(L.y), PK_F(L.sp[0]), PK_F(L.sp[2]), PK_F(L.sp[3])));
    stcg32(cp + 32, make_int4(PK_F(L.sc[0]), PK_F(L.sc[3]), PK_F(L.sc[4]), PK_F(L.sc[6])),
           make_int4(PK_F(L.sc[7]), PK_F(L.sc[8]), L.id, L.meta));
#undef PK_F
}

template <>
__device__ __forceinline__ void store_landmark<double>(unsigned char* block, int capacity, int j, const Landmark& L,
                                                       unsigned old_key) {
    unsigned char* cp = const_cast<unsigned char*>(cold_ptr<double>(block, capacity, j));
    const unsigned key = color_key(L.r, L.g, L.b);
    if (key != old_key) __stcg(reinterpret_cast<unsigned*>(block) + j, key);
    stcg32(cp, mk_i4(L.r, L.g), mk_i4(L.b, L.x));
    stcg32(cp + 32, mk_i4(L.y, L.sp[0]), mk_i4(L.sp[1], L.sp[2]));
    stcg32(cp + 64, mk_i4(L.sp[3], L.sc[0]), mk_i4(L.sc[1], L.sc[2]));
    stcg32(cp + 96, mk_i4(L.sc[3], L.sc[4]), mk_i4(L.sc[5], L.sc[6]));
    stcg32(cp + 128, mk_i4(L.sc[7], L.sc[8]), make_int4(L.id, L.meta, 0, 0));
}

// ---- fp32 working copy <-> the 64-byte f32 record: plain bit moves, no conversions -----------------
__device__ __forceinline__ void decode_cold_f(const int4* c, LandmarkF& L) {
    L.r = __int_as_float(c[0].x);
    L.g = __int_as_float(c[0].y);
    L.b = __int_as_float(c[0].z);
    L.x = __int_as_float(c[0].w);
    L.y = __int_as_float(c[1].x);
    L.sp[0] = __int_as_float(c[1].y);
    L.sp[1] = __int_as_float(c[1].z);
    L.sp[2] = __int_as_float(c[1].w);
    L.sc[0] = __int_as_float(c[2].x);
    L.sc[1] = __int_as_float(c[2].y);
    L.sc[2] = __int_as_float(c[2].z);
    L.sc[3] = __int_as_float(c[2].w);
    L.sc[4] = __int_as_float(c[3].x);
    L.sc[5] = __int_as_float(c[3].y);
    L.id = c[3].z;
    L.meta = c[3].w;
}
template <typename T>
__device__ __forceinline__ void load_landmark(const unsigned char* block, int capacity, int j, LandmarkF& L) {
    static_assert(sizeof(typename Rec<T>::Cold) == 64, "fp32 arithmetic needs fp32 storage");
    const unsigned char* cp = cold_ptr<T>(block, capacity, j);
    int4 c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = ldcg16_rec(cp + 16 * i);
    decode_cold_f(c, L);
}
__device__ __forceinline__ unsigned color_key_f(float r, float g, float b) {
    const unsigned kr = (unsigned)__float2int_rn(fminf(fmaxf(r, 0.0f), 255.0f));
    const unsigned kg = (unsigned)__float2int_rn(fminf(fmaxf(g, 0.0f), 255.0f));
    const unsigned kb = (unsigned)__float2int_rn(fminf(fmaxf(b, 0.0f), 255.0f));
    return kr | (kg << 8) | (kb << 16);
}
template <typename T>
__device__ __forceinline__ void store_landmark(unsigned char* block, int capacity, int j, const LandmarkF& L,
                                               unsigned old_key = kNoKey) {
    static_assert(sizeof(typename Rec<T>::Cold) == 64, "fp32 arithmetic needs fp32 storage");
    unsigned char* cp = const_cast<unsigned char*>(cold_ptr<T>(block, capacity, j));
#define PK_FI(v) __float_as_int(v)
    const unsigned key = color_key_f(L.r, L.g, L.b);
    if (key != old_key) __stcg(reinterpret_cast<unsigned*>(block) + j, key);
    stcg32(cp, make_int4(PK_FI(L.r), PK_FI(L.g), PK_FI(L.b), PK_FI(L.x)),
           make_int4(PK_FI(L.y), PK_FI(L.sp[0]), PK_FI(L.sp[1]), PK_FI(L.sp[2])));
    stcg32(cp + 32, make_int4(PK_FI(L.sc[0]), PK_FI(L.sc[1]), PK_FI(L.sc[2]), PK_FI(L.sc[3])),
           make_int4(PK_FI(L.sc[4]), PK_FI(L.sc[5]), L.id, L.meta));
#undef PK_FI
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copies (cp.async.bulk; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, bulk-group completion
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
// variants taking precomputed 32-bit shared-window addresses (no generic->shared conversion)
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, unsigned parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d_a(uint32_t dst, const void* src_gmem, unsigned bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
// per-thread 16-byte asynchronous copy global -> shared (LDGSTS), L2-only caching; completion by
// commit/wait groups.  Unlike cp.async.bulk (UBLKCP, uniform operands) every lane can use its own
// addresses in one instruction, which is what a scattered per-lane record fetch needs.
__device__ __forceinline__ void cp_async16_a(uint32_t dst, const void* src_gmem) {
    // .L2::64B: fetch only the record's 64-byte half of the 128-byte line on an L2 miss (see ldcg16_rec)
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(dst), "l"(src_gmem) : "memory");
}
// same for streams that cover whole 128-byte lines anyway (keys, poses): default L2 fetch size
__device__ __forceinline__ void cp_async16_line(uint32_t dst, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// sum of the absolute differences of the four byte lanes, plus c (one VABSDIFF4.U8.ACC)
__device__ __forceinline__ unsigned vsad4_acc(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int4 lds16_a(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// double-double arithmetic (error-free transformations) for the weight prefix
// ---------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd two_sum(double a, double b) {
    double s = __dadd_rn(a, b);
    double bb = __dsub_rn(s, a);
    double e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    return dd{s, e};
}
__device__ __forceinline__ dd quick_two_sum(double a, double b) {
    double s = __dadd_rn(a, b);
    double e = __dsub_rn(b, __dsub_rn(s, a));
    return dd{s, e};
}
__device__ __forceinline__ dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo = __dadd_rn(s.lo, a.lo);
    return quick_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo = __dadd_rn(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = __dadd_rn(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}

template <typename T>
__device__ __forceinline__ void load_staged(uint32_t rec, Landmark& L) {
    constexpr int kWords = (int)(sizeof(typename Rec<T>::Cold) / 16);
    int4 c[kWords];
#pragma unroll
    for (int i = 0; i < kWords; ++i) c[i] = lds16_a(rec + 16 * i);
    decode_cold<T>(c, L);
}

// key the record currently has in the hot region (always derived from the stored colour)
__device__ __forceinline__ unsigned stored_key(const Landmark& L, int dtype_tag) {
    if (dtype_tag == PK_DTYPE_F32) return color_key((double)(float)L.r, (double)(float)L.g, (double)(float)L.b);
    return color_key(L.r, L.g, L.b);
}
__device__ __forceinline__ unsigned stored_key(const LandmarkF& L, int) { return color_key_f(L.r, L.g, L.b); }

template <typename T>
__device__ __forceinline__ void load_staged(uint32_t rec, LandmarkF& L) {
    static_assert(sizeof(typename Rec<T>::Cold) == 64, "fp32 arithmetic needs fp32 storage");
    int4 c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = lds16_a(rec + 16 * i);
    decode_cold_f(c, L);
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
#endif  // __CUDACC__

}  // namespace pk
