// Probe entry points: the per-landmark math of the fused kernel exposed one triple per thread.
// They back the scalar helper methods of the reference's FilterParticle
// (probability_of_match prkt_core_v2.py:383-455; generate_measurement / measurement_jacobian /
// measurement_covariance / kalman_gain / importance_factor / Feature.update_* :748-930) on the
// device, and let the parity tests pin the device arithmetic against the reference's
// known-answer vectors without going through a whole filter frame.
#include "pk_common.cuh"
#include "pk_filter_math.cuh"

namespace pk {

__device__ __forceinline__ void gather_lm(Landmark& L, const double* mean5, const double* covp, const double* covc,
                                          long long i) {
    L.x = mean5[5 * i + 0];
    L.y = mean5[5 * i + 1];
    L.r = mean5[5 * i + 2];
    L.g = mean5[5 * i + 3];
    L.b = mean5[5 * i + 4];
#pragma unroll
    for (int q = 0; q < 4; ++q) L.sp[q] = covp[4 * i + q];
#pragma unroll
    for (int q = 0; q < 9; ++q) L.sc[q] = covc[9 * i + q];
}

__global__ void probe_likelihood_kernel(const double* __restrict__ pose3, const double* __restrict__ blob4,
                                        const double* __restrict__ dir2, const double* __restrict__ mean5,
                                        const double* __restrict__ covp, const double* __restrict__ covc, long long n,
                                        pk_params prm, double* __restrict__ out, unsigned* __restrict__ flags_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Landmark L;
    gather_lm(L, mean5, covp, covc, i);
    L.meta = 0;
    L.id = 1;
    unsigned flags = 0;
    double pse;
    out[i] = match_likelihood(L, pose3[3 * i], pose3[3 * i + 1], pose3[3 * i + 2], blob4[4 * i], blob4[4 * i + 1],
                              blob4[4 * i + 2], blob4[4 * i + 3], dir2[2 * i], dir2[2 * i + 1], prm, flags, pse);
    if (flags && flags_out) atomicOr(flags_out, flags);
}

__global__ void probe_ekf_kernel(const double* __restrict__ pose2, const double* __restrict__ blob4,
                                 const double* __restrict__ mean5, const double* __restrict__ covp,
                                 const double* __restrict__ covc, const int* __restrict__ meta, long long n, pk_params prm,
                                 double* __restrict__ mean5_out, double* __restrict__ covp_out,
                                 double* __restrict__ covc_out, double* __restrict__ factor_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Landmark L;
    gather_lm(L, mean5, covp, covc, i);
    L.meta = meta ? meta[i] : 0;
    L.id = (L.meta & PK_META_POTENTIAL) ? -1 : 1;
    unsigned flags = 0;
    int id_out = 0, promoted = 0;
    bool changed = false;
    factor_out[i] = ekf_update_lm(L, pose2[2 * i], pose2[2 * i + 1], blob4[4 * i], blob4[4 * i + 1], blob4[4 * i + 2],
                                  blob4[4 * i + 3], prm, id_out, flags, promoted, changed);
    mean5_out[5 * i + 0] = L.x;
    mean5_out[5 * i + 1] = L.y;
    mean5_out[5 * i + 2] = L.r;
    mean5_out[5 * i + 3] = L.g;
    mean5_out[5 * i + 4] = L.b;
#pragma unroll
    for (int q = 0; q < 4; ++q) covp_out[4 * i + q] = L.sp[q];
#pragma unroll
    for (int q = 0; q < 9; ++q) covc_out[9 * i + q] = L.sc[q];
}

}  // namespace pk

using namespace pk;

extern "C" {

int pk_probe_likelihood(const double* pose3, const double* blob4, const double* dir2, const double* mean5,
                        const double* covp, const double* covc, long long n, const pk_params* params, double* out,
                        void* stream) {
    PK_CHECK_ARG(pose3 && blob4 && dir2 && mean5 && covp && covc && out && params, "null pointer");
    PK_CHECK_ARG(n >= 0, "n < 0");
    if (n == 0) return PK_OK;
    probe_likelihood_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pose3, blob4, dir2, mean5, covp,
                                                                                         covc, n, *params, out, nullptr);
    PK_LAUNCH_CHECK("probe_likelihood_kernel");
    return PK_OK;
}

int pk_probe_ekf(const double* pose2, const double* blob4, const double* mean5, const double* covp, const double* covc,
                 const int* meta, long long n, const pk_params* params, double* mean5_out, double* covp_out,
                 double* covc_out, double* factor_out, void* stream) {
    PK_CHECK_ARG(pose2 && blob4 && mean5 && covp && covc && params && mean5_out && covp_out && covc_out && factor_out,
                 "null pointer");
    PK_CHECK_ARG(n >= 0, "n < 0");
    if (n == 0) return PK_OK;
    probe_ekf_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        pose2, blob4, mean5, covp, covc, meta, n, *params, mean5_out, covp_out, covc_out, factor_out);
    PK_LAUNCH_CHECK("probe_ekf_kernel");
    return PK_OK;
}

}  // extern "C"
