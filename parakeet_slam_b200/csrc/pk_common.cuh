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
// Landmark record layout.  One particle's map ("block") = [hot x capacity][cold x capacity].
// The hot part is all the association pre-filter needs (colour mean + meta) and is what the
// fused kernel streams through shared memory; the cold part (position, covariance blocks, id) is
// fetched only for the few landmarks that survive the colour gate.  A cold record is exactly two
// (f32) / four (f64) 32-byte DRAM sectors.
// ---------------------------------------------------------------------------------------------
struct alignas(16) HotF {
    float r, g, b;
    int meta;
};
struct alignas(16) ColdF {
    float x, y;
    float sp[4];  // position covariance block, row-major [[a,b],[c,d]]
    float sc[9];  // colour covariance block, row-major
    int id;       // reference landmark id (>0 full, <0 potential)
};
struct alignas(16) HotD {
    double r, g, b;
    int meta;
    int pad;
};
struct alignas(16) ColdD {
    double x, y;
    double sp[4];
    double sc[9];
    int id;
    int pad;
};
static_assert(sizeof(HotF) == 16 && sizeof(ColdF) == 64, "f32 record layout");
static_assert(sizeof(HotD) == 32 && sizeof(ColdD) == 128, "f64 record layout");

template <typename T>
struct Rec;
template <>
struct Rec<float> {
    using Hot = HotF;
    using Cold = ColdF;
    static constexpr int kDtype = PK_DTYPE_F32;
};
template <>
struct Rec<double> {
    using Hot = HotD;
    using Cold = ColdD;
    static constexpr int kDtype = PK_DTYPE_F64;
};

// fp64 working copy of one landmark
struct Landmark {
    double x, y, r, g, b;
    double sp[4];
    double sc[9];
    int meta;
    int id;
};

__host__ __device__ inline size_t hot_bytes(int dtype) { return dtype == PK_DTYPE_F64 ? sizeof(HotD) : sizeof(HotF); }
__host__ __device__ inline size_t cold_bytes(int dtype) { return dtype == PK_DTYPE_F64 ? sizeof(ColdD) : sizeof(ColdF); }
__host__ __device__ inline size_t block_bytes(int capacity, int dtype) {
    return (size_t)capacity * (hot_bytes(dtype) + cold_bytes(dtype));
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Cache-global (L2) vector loads/stores: landmark records are written and re-read by different
// lanes of a warp inside one kernel, so they must never be served from a stale L1 line.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int4 ldcg16(const void* p) { return __ldcg(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ void stcg16(void* p, int4 v) { __stcg(reinterpret_cast<int4*>(p), v); }

template <typename T>
__device__ __forceinline__ void load_landmark(const unsigned char* block, int capacity, int j, Landmark& L);

template <>
__device__ __forceinline__ void load_landmark<float>(const unsigned char* block, int capacity, int j, Landmark& L) {
    const unsigned char* hp = block + (size_t)j * sizeof(HotF);
    const unsigned char* cp = block + (size_t)capacity * sizeof(HotF) + (size_t)j * sizeof(ColdF);
    int4 h = ldcg16(hp);
    int4 c0 = ldcg16(cp), c1 = ldcg16(cp + 16), c2 = ldcg16(cp + 32), c3 = ldcg16(cp + 48);
    L.r = (double)__int_as_float(h.x);
    L.g = (double)__int_as_float(h.y);
    L.b = (double)__int_as_float(h.z);
    L.meta = h.w;
    L.x = (double)__int_as_float(c0.x);
    L.y = (double)__int_as_float(c0.y);
    L.sp[0] = (double)__int_as_float(c0.z);
    L.sp[1] = (double)__int_as_float(c0.w);
    L.sp[2] = (double)__int_as_float(c1.x);
    L.sp[3] = (double)__int_as_float(c1.y);
    L.sc[0] = (double)__int_as_float(c1.z);
    L.sc[1] = (double)__int_as_float(c1.w);
    L.sc[2] = (double)__int_as_float(c2.x);
    L.sc[3] = (double)__int_as_float(c2.y);
    L.sc[4] = (double)__int_as_float(c2.z);
    L.sc[5] = (double)__int_as_float(c2.w);
    L.sc[6] = (double)__int_as_float(c3.x);
    L.sc[7] = (double)__int_as_float(c3.y);
    L.sc[8] = (double)__int_as_float(c3.z);
    L.id = c3.w;
}

__device__ __forceinline__ double i4lo(int4 v) { return __hiloint2double(v.y, v.x); }
__device__ __forceinline__ double i4hi(int4 v) { return __hiloint2double(v.w, v.z); }
__device__ __forceinline__ int4 mk_i4(double a, double b) {
    return make_int4(__double2loint(a), __double2hiint(a), __double2loint(b), __double2hiint(b));
}

template <>
__device__ __forceinline__ void load_landmark<double>(const unsigned char* block, int capacity, int j, Landmark& L) {
    const unsigned char* hp = block + (size_t)j * sizeof(HotD);
    const unsigned char* cp = block + (size_t)capacity * sizeof(HotD) + (size_t)j * sizeof(ColdD);
    int4 h0 = ldcg16(hp), h1 = ldcg16(hp + 16);
    L.r = i4lo(h0);
    L.g = i4hi(h0);
    L.b = i4lo(h1);
    L.meta = h1.z;
    int4 c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = ldcg16(cp + 16 * i);
    L.x = i4lo(c[0]);
    L.y = i4hi(c[0]);
    L.sp[0] = i4lo(c[1]);
    L.sp[1] = i4hi(c[1]);
    L.sp[2] = i4lo(c[2]);
    L.sp[3] = i4hi(c[2]);
    L.sc[0] = i4lo(c[3]);
    L.sc[1] = i4hi(c[3]);
    L.sc[2] = i4lo(c[4]);
    L.sc[3] = i4hi(c[4]);
    L.sc[4] = i4lo(c[5]);
    L.sc[5] = i4hi(c[5]);
    L.sc[6] = i4lo(c[6]);
    L.sc[7] = i4hi(c[6]);
    L.sc[8] = i4lo(c[7]);
    L.id = c[7].z;
}

template <typename T>
__device__ __forceinline__ void store_landmark(unsigned char* block, int capacity, int j, const Landmark& L);

template <>
__device__ __forceinline__ void store_landmark<float>(unsigned char* block, int capacity, int j, const Landmark& L) {
    unsigned char* hp = block + (size_t)j * sizeof(HotF);
    unsigned char* cp = block + (size_t)capacity * sizeof(HotF) + (size_t)j * sizeof(ColdF);
#define PK_F(v) __float_as_int((float)(v))
    stcg16(hp, make_int4(PK_F(L.r), PK_F(L.g), PK_F(L.b), L.meta));
    stcg16(cp, make_int4(PK_F(L.x), PK_F(L.y), PK_F(L.sp[0]), PK_F(L.sp[1])));
    stcg16(cp + 16, make_int4(PK_F(L.sp[2]), PK_F(L.sp[3]), PK_F(L.sc[0]), PK_F(L.sc[1])));
    stcg16(cp + 32, make_int4(PK_F(L.sc[2]), PK_F(L.sc[3]), PK_F(L.sc[4]), PK_F(L.sc[5])));
    stcg16(cp + 48, make_int4(PK_F(L.sc[6]), PK_F(L.sc[7]), PK_F(L.sc[8]), L.id));
#undef PK_F
}

template <>
__device__ __forceinline__ void store_landmark<double>(unsigned char* block, int capacity, int j, const Landmark& L) {
    unsigned char* hp = block + (size_t)j * sizeof(HotD);
    unsigned char* cp = block + (size_t)capacity * sizeof(HotD) + (size_t)j * sizeof(ColdD);
    stcg16(hp, mk_i4(L.r, L.g));
    int4 h1 = mk_i4(L.b, 0.0);
    h1.z = L.meta;
    h1.w = 0;
    stcg16(hp + 16, h1);
    stcg16(cp, mk_i4(L.x, L.y));
    stcg16(cp + 16, mk_i4(L.sp[0], L.sp[1]));
    stcg16(cp + 32, mk_i4(L.sp[2], L.sp[3]));
    stcg16(cp + 48, mk_i4(L.sc[0], L.sc[1]));
    stcg16(cp + 64, mk_i4(L.sc[2], L.sc[3]));
    stcg16(cp + 80, mk_i4(L.sc[4], L.sc[5]));
    stcg16(cp + 96, mk_i4(L.sc[6], L.sc[7]));
    int4 t = mk_i4(L.sc[8], 0.0);
    t.z = L.id;
    t.w = 0;
    stcg16(cp + 112, t);
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

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
#endif  // __CUDACC__

}  // namespace pk
