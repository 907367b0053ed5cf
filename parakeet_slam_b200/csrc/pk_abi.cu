// C ABI plumbing: errors, layout queries, construction and map import/export kernels.
// Replaces FastSLAM.__init__ / FilterParticle.__init__ / load_feature_list
// (reference prkt_core_v2.py:38-57, 279-299) as bulk device initialisation.
#include <stdarg.h>
#include <stdio.h>

#include "pk_common.cuh"

namespace pk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return PK_ECUDA;
}

int num_sms() {
    // per device (a process may drive filters on several GPUs) and cheap enough to ask every time it matters
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        cached = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
        cached_dev = dev;
    }
    return cached;
}

// ---------------------------------------------------------------------------------------------
__global__ void init_particles_kernel(double* __restrict__ pose4, int* __restrict__ slot, int* __restrict__ aux2,
                                      long long M, int n_live, int next_id) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double2* p = reinterpret_cast<double2*>(pose4 + 4 * i);
    p[0] = make_double2(0.0, 0.0);  // x, y               FilterParticle.__init__ :282-283
    p[1] = make_double2(0.0, 1.0);  // heading 0 (:284), weight 1 (:288)
    slot[i] = (int)i;
    reinterpret_cast<int2*>(aux2)[i] = make_int2(n_live, next_id);
}

template <typename T>
__global__ void map_broadcast_kernel(unsigned char* __restrict__ pool, int capacity, size_t bbytes, long long slot_lo, long long n_slots,
                                     int n, const double* __restrict__ mean5, const double* __restrict__ covp,
                                     const double* __restrict__ covc, const int* __restrict__ meta,
                                     const int* __restrict__ ids) {
    // one thread per (slot, landmark)
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_slots * n) return;
    long long s = slot_lo + t / n;
    int j = (int)(t % n);
    Landmark L;
    L.x = mean5[5 * j + 0];
    L.y = mean5[5 * j + 1];
    L.r = mean5[5 * j + 2];
    L.g = mean5[5 * j + 3];
    L.b = mean5[5 * j + 4];
#pragma unroll
    for (int q = 0; q < 4; ++q) L.sp[q] = covp[4 * j + q];
#pragma unroll
    for (int q = 0; q < 9; ++q) L.sc[q] = covc[9 * j + q];
    L.meta = meta[j];
    L.id = ids[j];
    unsigned char* block = pool + (size_t)s * bbytes;
    store_landmark<T>(block, capacity, j, L);
}

template <typename T, bool kImport>
__global__ void map_xfer_kernel(unsigned char* __restrict__ pool, int capacity, size_t bbytes, const int* __restrict__ slot,
                                long long p_lo, long long count, double* __restrict__ mean5, double* __restrict__ covp,
                                double* __restrict__ covc, int* __restrict__ meta, int* __restrict__ ids) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * capacity) return;
    long long pi = t / capacity;
    int j = (int)(t % capacity);
    unsigned char* block = pool + (size_t)slot[p_lo + pi] * bbytes;
    Landmark L;
    if (kImport) {
        L.x = mean5[5 * t + 0];
        L.y = mean5[5 * t + 1];
        L.r = mean5[5 * t + 2];
        L.g = mean5[5 * t + 3];
        L.b = mean5[5 * t + 4];
#pragma unroll
        for (int q = 0; q < 4; ++q) L.sp[q] = covp[4 * t + q];
#pragma unroll
        for (int q = 0; q < 9; ++q) L.sc[q] = covc[9 * t + q];
        L.meta = meta[t];
        L.id = ids[t];
        store_landmark<T>(block, capacity, j, L);
    } else {
        load_landmark<T>(block, capacity, j, L);
        mean5[5 * t + 0] = L.x;
        mean5[5 * t + 1] = L.y;
        mean5[5 * t + 2] = L.r;
        mean5[5 * t + 3] = L.g;
        mean5[5 * t + 4] = L.b;
#pragma unroll
        for (int q = 0; q < 4; ++q) covp[4 * t + q] = L.sp[q];
#pragma unroll
        for (int q = 0; q < 9; ++q) covc[9 * t + q] = L.sc[q];
        meta[t] = L.meta;
        ids[t] = L.id;
    }
}

}  // namespace pk

using namespace pk;

extern "C" {

int pk_version(void) { return PK_ABI_VERSION; }

const char* pk_last_error(void) { return g_err; }

int pk_default_params(pk_params* out) {
    PK_CHECK_ARG(out != nullptr, "out is NULL");
    out->bearing_gate = 0.5;                       // prkt_core_v2.py:433
    out->position_gate = 3.141592653589793 / 2.0;  // prkt_core_v2.py:474  math.pi/2
    out->color_gate = 300.0;                       // prkt_core_v2.py:441
    out->no_match_weight = 0.1;                    // prkt_core_v2.py:857
    out->qt_diag = 0.1;                            // prkt_core_v2.py:50-53
    out->promote_count = 5;                        // prkt_core_v2.py:114
    out->model = 0;                                // the reference's measurement model, as written
    return PK_OK;
}

int pk_check_device(void) {
    int dev = 0;
    PK_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    PK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    PK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) {
        set_error("device compute capability %d.%d is not sm_100 (this library is built for sm_100a only)", major, minor);
        return PK_EARCH;
    }
    return PK_OK;
}

int pk_hot_bytes(int dtype) { return (int)hot_bytes(dtype); }
int pk_cold_bytes(int dtype) { return (int)cold_bytes(dtype); }
long long pk_block_bytes(int capacity, int dtype) { return (long long)block_bytes(capacity, dtype); }

int pk_init_particles(double* pose4, int* slot, int* aux2, long long M, int n_live, int next_id, void* stream) {
    PK_CHECK_ARG(pose4 && slot && aux2, "null pointer");
    PK_CHECK_ARG(M >= 0, "M < 0");
    if (M == 0) return PK_OK;
    const int threads = 256;
    long long blocks = (M + threads - 1) / threads;
    init_particles_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(pose4, slot, aux2, M, n_live, next_id);
    PK_LAUNCH_CHECK("init_particles_kernel");
    return PK_OK;
}

int pk_map_broadcast(void* pool, int capacity, int dtype, long long slot_lo, long long slot_hi, int n,
                     const double* mean5, const double* covp, const double* covc, const int* meta, const int* ids,
                     void* stream) {
    PK_CHECK_ARG(pool != nullptr, "pool is NULL");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    PK_CHECK_ARG(n >= 0 && n <= capacity, "n out of range");
    PK_CHECK_ARG(slot_hi >= slot_lo && slot_lo >= 0, "slot range");
    if (n == 0 || slot_hi == slot_lo) return PK_OK;
    PK_CHECK_ARG(mean5 && covp && covc && meta && ids, "null map pointer");
    long long total = (slot_hi - slot_lo) * n;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    PK_CHECK_ARG(blocks < (1ll << 31), "too many blocks");
    const size_t bb = block_bytes(capacity, dtype);
    if (dtype_base(dtype) == PK_DTYPE_F32)
        map_broadcast_kernel<float><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            (unsigned char*)pool, capacity, bb, slot_lo, slot_hi - slot_lo, n, mean5, covp, covc, meta, ids);
    else
        map_broadcast_kernel<double><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            (unsigned char*)pool, capacity, bb, slot_lo, slot_hi - slot_lo, n, mean5, covp, covc, meta, ids);
    PK_LAUNCH_CHECK("map_broadcast_kernel");
    return PK_OK;
}

static int map_xfer(bool import, void* pool, int capacity, int dtype, const int* slot, long long p_lo, long long count,
                    double* mean5, double* covp, double* covc, int* meta, int* ids, void* stream) {
    PK_CHECK_ARG(pool && slot && mean5 && covp && covc && meta && ids, "null pointer");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    PK_CHECK_ARG(capacity > 0 && count >= 0 && p_lo >= 0, "sizes");
    if (count == 0) return PK_OK;
    long long total = count * capacity;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    PK_CHECK_ARG(blocks < (1ll << 31), "too many blocks");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* pl = (unsigned char*)pool;
    const size_t bb = block_bytes(capacity, dtype);
    if (dtype_base(dtype) == PK_DTYPE_F32) {
        if (import)
            map_xfer_kernel<float, true><<<(unsigned)blocks, threads, 0, st>>>(pl, capacity, bb, slot, p_lo, count, mean5, covp, covc, meta, ids);
        else
            map_xfer_kernel<float, false><<<(unsigned)blocks, threads, 0, st>>>(pl, capacity, bb, slot, p_lo, count, mean5, covp, covc, meta, ids);
    } else {
        if (import)
            map_xfer_kernel<double, true><<<(unsigned)blocks, threads, 0, st>>>(pl, capacity, bb, slot, p_lo, count, mean5, covp, covc, meta, ids);
        else
            map_xfer_kernel<double, false><<<(unsigned)blocks, threads, 0, st>>>(pl, capacity, bb, slot, p_lo, count, mean5, covp, covc, meta, ids);
    }
    PK_LAUNCH_CHECK("map_xfer_kernel");
    return PK_OK;
}

int pk_map_export(const void* pool, int capacity, int dtype, const int* slot, long long p_lo, long long count,
                  double* mean5, double* covp, double* covc, int* meta, int* ids, void* stream) {
    return map_xfer(false, const_cast<void*>(pool), capacity, dtype, slot, p_lo, count, mean5, covp, covc, meta, ids, stream);
}

int pk_map_import(void* pool, int capacity, int dtype, const int* slot, long long p_lo, long long count,
                  const double* mean5, const double* covp, const double* covc, const int* meta, const int* ids,
                  void* stream) {
    return map_xfer(true, pool, capacity, dtype, slot, p_lo, count, const_cast<double*>(mean5), const_cast<double*>(covp),
                    const_cast<double*>(covc), const_cast<int*>(meta), const_cast<int*>(ids), stream);
}

}  // extern "C"
