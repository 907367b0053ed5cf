// K1 -- motion sampling.  Replaces FastSLAM.motion_update / motion_model
// (reference prkt_core_v2.py:148-208) and the heading<->quaternion round trip of utils.py:8-35
// (tf.transformations quaternion_from_euler / euler_from_quaternion, 'sxyz').
//
// One thread per particle, pose records are 32-byte (x, y, heading, weight) so a warp reads and
// writes one contiguous kilobyte.  Arithmetic that decides x and y is written with explicit
// round-to-nearest intrinsics (no FMA contraction) so that, with injected noise, positions
// reproduce the reference's Python-float arithmetic operation for operation.
#include "pk_common.cuh"
#include "pk_filter_math.cuh"

namespace pk {

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        uint64_t p0 = (uint64_t)M0 * c[0];
        uint64_t p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0;
        uint32_t n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = lo1;
        c[2] = n2;
        c[3] = lo0;
    }
    __host__ __device__ static inline void run(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += W0;
            k1 += W1;
        }
    }
};

__device__ __forceinline__ float u32_open(uint32_t x) {
    // (0,1) uniform from 32 random bits, never 0 or 1: the 24 leading bits plus one half (Box-Muller tails reach 5.8 sigma)
    return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// three standard normals from ONE Philox4x32-10 call: its four words are the uniforms of two Box-Muller pairs.
// The transform runs in fp32 (a relative 1e-7 on a noise sample that is scaled by ~0.01 m / rad is far below the
// sample's own spread; the kernel is bound by its fp64 instructions, not by memory); poses stay fp64.
__device__ __forceinline__ void philox_normals3(unsigned long long seed, unsigned long long frame,
                                                unsigned long long particle, double& z0, double& z1, double& z2) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t a[4] = {(uint32_t)particle, (uint32_t)(particle >> 32), (uint32_t)frame, (uint32_t)(frame >> 32)};
    Philox::run(a, k0, k1);
    const float r1 = sqrtf(-2.0f * logf(u32_open(a[0])));
    const float r2 = sqrtf(-2.0f * logf(u32_open(a[2])));
    float s, c;
    sincospif(2.0f * u32_open(a[1]), &s, &c);
    z0 = (double)(r1 * c);
    z1 = (double)(r1 * s);
    sincospif(2.0f * u32_open(a[3]), &s, &c);
    z2 = (double)(r2 * c);
}

// heading -> quaternion (0,0,sin h/2,cos h/2) -> heading, operation for operation as
// tf.transformations does it (quaternion_matrix + euler_from_matrix 'sxyz'); see
// parakeet_slam_b200/rosless/transformations.py for the restated algorithm.
__device__ __forceinline__ double wrap_heading(double h) {
    double half = h / 2.0;
    double z, w;
    sincos(half, &z, &w);
    double nq = __dadd_rn(__dmul_rn(z, z), __dmul_rn(w, w));
    double s = sqrt(2.0 / nq);
    double zs = __dmul_rn(z, s), ws = __dmul_rn(w, s);
    double m10 = __dmul_rn(zs, ws);
    double m00 = __dsub_rn(1.0, __dmul_rn(zs, zs));
    return pk_atan2(m10, m00);   // branch-free, <= 1.5 ulp (libm's atan2 costs several times more here)
}

__global__ void __launch_bounds__(256)
motion_kernel(double* __restrict__ pose4, long long M, const double* __restrict__ noise3, unsigned long long seed,
              unsigned long long frame, long long particle_offset, double vdt, double half_dheading, double sd, double sh) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double2* rec = reinterpret_cast<double2*>(pose4 + 4 * i);
    double2 xy = rec[0];
    double2 tw = rec[1];
    double z0, z1, z2;
    if (noise3 != nullptr) {
        z0 = noise3[3 * i + 0];
        z1 = noise3[3 * i + 1];
        z2 = noise3[3 * i + 2];
    } else {
        philox_normals3(seed, frame, (unsigned long long)(particle_offset + i), z0, z1, z2);
    }
    // prkt_core_v2.py:185-194
    double ds = __dadd_rn(vdt, __dmul_rn(sd, z0));
    double h1 = __dadd_rn(__dadd_rn(tw.x, half_dheading), __dmul_rn(sh, z1));
    double h2 = __dadd_rn(__dadd_rn(h1, half_dheading), __dmul_rn(sh, z2));
    double s1, c1;
    sincos(h1, &s1, &c1);
    // :198-204
    xy.x = __dadd_rn(xy.x, __dmul_rn(ds, c1));
    xy.y = __dadd_rn(xy.y, __dmul_rn(ds, s1));
    // :206 heading_to_quaternion, then every later read goes through quaternion_to_heading
    tw.x = wrap_heading(h2);
    rec[0] = xy;
    rec[1] = tw;
}

}  // namespace pk

using namespace pk;

extern "C" int pk_motion_update(double* pose4, long long M, const double* noise3, unsigned long long seed,
                                unsigned long long frame, long long particle_offset, double v, double w, double dt,
                                void* stream) {
    PK_CHECK_ARG(pose4 != nullptr, "pose4 is NULL");
    PK_CHECK_ARG(M >= 0, "M < 0");
    if (M == 0) return PK_OK;
    // host side of motion_model: scalars shared by all particles (prkt_core_v2.py:183-193)
    volatile double sd = fabs(.05 * v) + fabs(.005 * w) + .0005;
    volatile double sh = fabs(.025 * w) + fabs(.005 * v) + .0005;
    volatile double dheading = w * dt;
    volatile double half = dheading / 2;
    volatile double vdt = v * dt;
    const int threads = 256;
    long long blocks = (M + threads - 1) / threads;
    PK_CHECK_ARG(blocks < (1ll << 31), "too many blocks");
    motion_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(pose4, M, noise3, seed, frame, particle_offset,
                                                                          vdt, half, sd, sh);
    PK_LAUNCH_CHECK("motion_kernel");
    return PK_OK;
}
