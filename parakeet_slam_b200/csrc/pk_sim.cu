// Device-side stand-ins for the two tools around the filter:
//
//  * scan simulator -- what the un-vendored `viz_feature_sim` node feeds the reference: per frame a
//    VizScan of K blobs (bearing, r, g, b) (`matrix.py:35-39`, `prkt_core_v2.py:344`).  The rule is the
//    `synth360` scenario's (parakeet_slam_b200/scenario.py): the K landmarks nearest the
//    true pose in ascending distance order (stable), bearing = wrap_pi(atan2(ly-y, lx-x) - theta) +
//    N(0, sigma_b^2), colour = truth + N(0, sigma_c^2).  With the scan left on the device
//    (pk_measurement_update_dev) a long-horizon run never touches the host;
//  * accuracy analysis -- the working form of `analyze_slam.py:1-36` (squared x / y error of the
//    estimate against the truth) plus heading error (`utils.py:heading_error`, `minimize_angle`),
//    the weight statistics (sum w, sum w^2 -> N_eff) and per-landmark map error over all particles.
#include <math.h>

#include "pk_common.cuh"

namespace pk {

constexpr int kSimThreads = 1024;

__device__ __forceinline__ double wrap_pi_dev(double a) {
    // (a + pi) % (2 pi) - pi with Python's sign convention for %
    const double two_pi = 2.0 * 3.141592653589793;
    double r = fmod(__dadd_rn(a, 3.141592653589793), two_pi);
    if (r < 0.0) r = __dadd_rn(r, two_pi);
    return __dsub_rn(r, 3.141592653589793);
}

// Philox-free counter hash for the optional on-device observation noise (throughput runs only)
__device__ __forceinline__ double sim_normal(unsigned long long seed, unsigned long long frame, unsigned idx) {
    unsigned long long z = seed ^ (frame * 0x9E3779B97F4A7C15ull) ^ ((unsigned long long)idx * 0xBF58476D1CE4E5B9ull);
    auto mix = [](unsigned long long v) {
        v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ull; v ^= v >> 27; v *= 0x94D049BB133111EBull; v ^= v >> 31;
        return v;
    };
    const unsigned long long a = mix(z), b = mix(z + 0x9E3779B97F4A7C15ull);
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

__global__ void __launch_bounds__(kSimThreads)
simulate_scan_kernel(const double* __restrict__ lm5, int N, double x, double y, double th, int K,
                     const double* __restrict__ noise4, unsigned long long seed, unsigned long long frame, double sigma_b,
                     double sigma_c, double* __restrict__ d2_ws, double* __restrict__ obs, int* __restrict__ lm_idx) {
    __shared__ double w_d[32];
    __shared__ int w_i[32];
    __shared__ int chosen[PK_MAX_OBS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int j = t; j < N; j += kSimThreads) {
        const double dx = __dsub_rn(lm5[5 * j], x), dy = __dsub_rn(lm5[5 * j + 1], y);
        d2_ws[j] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));  // (lx-x)**2 + (ly-y)**2
    }
    __syncthreads();
    const int Keff = min(K, N);
    for (int k = 0; k < Keff; ++k) {
        // stable arg-min: smallest distance, lowest index on ties (numpy argsort kind="stable")
        double bd = INFINITY;
        int bi = 0x7fffffff;
        for (int j = t; j < N; j += kSimThreads) {
            const double d = d2_ws[j];
            if (d < bd || (d == bd && j < bi)) { bd = d; bi = j; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { w_d[warp] = bd; w_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bd = w_d[lane];
            bi = w_i[lane];
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (lane == 0) {
                chosen[k] = bi;
                d2_ws[bi] = INFINITY;  // taken (a NaN distance is never chosen, as in argsort it sorts last)
            }
        }
        __syncthreads();
    }
    if (t < K) {
        const int src = chosen[t < Keff ? t : Keff - 1];  // fewer landmarks than blobs: repeat the last one
        const int nrow = t < Keff ? t : Keff - 1;
        double z[4];
        for (int e = 0; e < 4; ++e)
            z[e] = noise4 ? noise4[4 * nrow + e] : sim_normal(seed, frame, (unsigned)(4 * nrow + e));
        const double lx = lm5[5 * src], ly = lm5[5 * src + 1];
        const double b = wrap_pi_dev(__dsub_rn(atan2(__dsub_rn(ly, y), __dsub_rn(lx, x)), th));
        obs[4 * t] = __dadd_rn(b, __dmul_rn(sigma_b, z[0]));
        obs[4 * t + 1] = __dadd_rn(lm5[5 * src + 2], __dmul_rn(sigma_c, z[1]));
        obs[4 * t + 2] = __dadd_rn(lm5[5 * src + 3], __dmul_rn(sigma_c, z[2]));
        obs[4 * t + 3] = __dadd_rn(lm5[5 * src + 4], __dmul_rn(sigma_c, z[3]));
        if (lm_idx) lm_idx[t] = src;
    }
}

// ---- accuracy -------------------------------------------------------------------------------------
constexpr int kAccBlocks = 512;
constexpr int kAccQ = 8;  // sum w, sum w^2, sum (x-xt)^2, sum (y-yt)^2, sum dtheta^2, sum x, sum y, max w

__device__ __forceinline__ double minimize_angle(double d) {  // utils.py:minimize_angle
    const double two_pi = 2.0 * 3.141592653589793, pi = 3.141592653589793;
    d = fmod(d, two_pi);
    if (d > pi) d -= two_pi;
    if (d < -pi) d += two_pi;
    return d;
}

__global__ void __launch_bounds__(256)
accuracy_partial_kernel(const double* __restrict__ pose4, long long M, double xt, double yt, double tt,
                        double* __restrict__ ws) {
    __shared__ double sh[kAccQ][8];
    double q[kAccQ] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
        const double2 xy = *reinterpret_cast<const double2*>(pose4 + 4 * i);
        const double2 tw = *reinterpret_cast<const double2*>(pose4 + 4 * i + 2);
        const double ex = xy.x - xt, ey = xy.y - yt, et = minimize_angle(tw.x - tt);
        q[0] += tw.y;
        q[1] += tw.y * tw.y;
        q[2] += ex * ex;
        q[3] += ey * ey;
        q[4] += et * et;
        q[5] += xy.x;
        q[6] += xy.y;
        q[7] = fmax(q[7], tw.y);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < kAccQ; ++e) {
        for (int o = 16; o > 0; o >>= 1) {
            const double v = __shfl_xor_sync(0xffffffffu, q[e], o);
            q[e] = (e == 7) ? fmax(q[e], v) : q[e] + v;
        }
        if (lane == 0) sh[e][warp] = q[e];
    }
    __syncthreads();
    if (threadIdx.x < kAccQ) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a = (threadIdx.x == 7) ? fmax(a, sh[threadIdx.x][w]) : a + sh[threadIdx.x][w];
        ws[threadIdx.x * kAccBlocks + blockIdx.x] = a;
    }
}

__global__ void accuracy_final_kernel(const double* __restrict__ ws, int nblocks, double* __restrict__ out) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q >= kAccQ) return;
    double a = 0.0;
    for (int b = lane; b < nblocks; b += 32) a = (q == 7) ? fmax(a, ws[q * kAccBlocks + b]) : a + ws[q * kAccBlocks + b];
    for (int o = 16; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(0xffffffffu, a, o);
        a = (q == 7) ? fmax(a, v) : a + v;
    }
    if (lane == 0) out[q] = a;
}

// per true landmark j: sum over particles of |mu - truth|^2 of the landmark carrying id j+1, and how many
// particles hold it (known-map ids, load_feature_list :294-299).  One thread per (particle, slot).
template <typename T>
__global__ void __launch_bounds__(256)
map_error_kernel(const unsigned char* __restrict__ pool, size_t bbytes, int capacity, const int* __restrict__ slot,
                 const int* __restrict__ aux2, long long M, const double* __restrict__ truth5, int N,
                 double* __restrict__ err2, unsigned long long* __restrict__ count) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * capacity) return;
    const long long p = t / capacity;
    const int j = (int)(t % capacity);
    if (j >= aux2[2 * p]) return;
    const unsigned char* block = pool + (size_t)slot[p] * bbytes;
    Landmark L;
    load_landmark<T>(block, capacity, j, L);
    const int id = L.id < 0 ? -L.id : L.id;
    if (id < 1 || id > N) return;
    const double dx = L.x - truth5[5 * (id - 1)], dy = L.y - truth5[5 * (id - 1) + 1];
    atomicAdd(&err2[id - 1], dx * dx + dy * dy);
    atomicAdd(&count[id - 1], 1ull);
}

}  // namespace pk

using namespace pk;

extern "C" {

int pk_simulate_scan(const double* landmarks5, int N, double x, double y, double theta, int K, const double* noise4,
                     unsigned long long seed, unsigned long long frame, double sigma_bearing, double sigma_color,
                     double* workspace, double* obs_out, int* landmark_out, void* stream) {
    PK_CHECK_ARG(landmarks5 && workspace && obs_out, "null pointer");
    PK_CHECK_ARG(N >= 1 && K >= 1 && K <= PK_MAX_OBS, "N >= 1 and 1 <= K <= PK_MAX_OBS");
    simulate_scan_kernel<<<1, kSimThreads, 0, (cudaStream_t)stream>>>(landmarks5, N, x, y, theta, K, noise4, seed, frame,
                                                                     sigma_bearing, sigma_color, workspace, obs_out,
                                                                     landmark_out);
    PK_LAUNCH_CHECK("simulate_scan_kernel");
    return PK_OK;
}

int pk_accuracy(const double* pose4, long long M, double x_true, double y_true, double theta_true, double* out8,
                double* workspace, void* stream) {
    PK_CHECK_ARG(pose4 && out8 && workspace, "null pointer");
    PK_CHECK_ARG(M > 0, "M <= 0");
    long long blocks = (M + 255) / 256;
    if (blocks > kAccBlocks) blocks = kAccBlocks;
    cudaStream_t st = (cudaStream_t)stream;
    accuracy_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(pose4, M, x_true, y_true, theta_true, workspace);
    PK_LAUNCH_CHECK("accuracy_partial_kernel");
    accuracy_final_kernel<<<1, kAccQ * 32, 0, st>>>(workspace, (int)blocks, out8);
    PK_LAUNCH_CHECK("accuracy_final_kernel");
    return PK_OK;
}

int pk_map_error(const void* pool, int capacity, int dtype, const int* slot, const int* aux2, long long M,
                 const double* truth5, int N, double* err2_out, unsigned long long* count_out, void* stream) {
    PK_CHECK_ARG(pool && slot && aux2 && truth5 && err2_out && count_out, "null pointer");
    PK_CHECK_ARG(dtype_valid(dtype), "dtype");
    PK_CHECK_ARG(M > 0 && capacity > 0 && N > 0, "sizes");
    cudaStream_t st = (cudaStream_t)stream;
    PK_CUDA(cudaMemsetAsync(err2_out, 0, (size_t)N * sizeof(double), st));
    PK_CUDA(cudaMemsetAsync(count_out, 0, (size_t)N * sizeof(unsigned long long), st));
    const long long total = M * capacity;
    const long long blocks = (total + 255) / 256;
    PK_CHECK_ARG(blocks < (1ll << 31), "too many blocks");
    const size_t bb = block_bytes(capacity, dtype);
    if (dtype_base(dtype) == PK_DTYPE_F32)
        map_error_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const unsigned char*)pool, bb, capacity, slot, aux2, M, truth5,
                                                                 N, err2_out, count_out);
    else
        map_error_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((const unsigned char*)pool, bb, capacity, slot, aux2, M,
                                                                  truth5, N, err2_out, count_out);
    PK_LAUNCH_CHECK("map_error_kernel");
    return PK_OK;
}

}  // extern "C"
