// What the memory system alone allows for the fused measurement kernel's access pattern (BASELINE config 2).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/k2_mem_probe tools/k2_mem_probe.cu
// Run:   tools/k2_mem_probe            (plain run, CUDA events; prints one line per pattern)
//
// Per particle K2 touches: the 32-byte pose record (read, 8 bytes of it written), 8 bytes of slot / aux, the 256-byte key
// row at the head of the particle's landmark block, eight 64-byte landmark records inside that 4.4 KB block (read, then
// written back), and 32 bytes of association ids.  1423 algorithmic bytes per particle; blocks are reached through a
// slot permutation (copy-on-resample scatters them).  This probe issues exactly those accesses with NO arithmetic in
// between, through a deep shared-memory ring (cp.async), with as many warps as fit -- an upper bound on what any
// implementation of K2 can reach with this layout.  Patterns:
//   all      everything above                     keys     key rows only            recs_r   record reads only
//   recs_rw  record reads + write-backs            stream   the same BYTES as `all`, but contiguous (plain copy)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cpa16_64(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st32(void* p, int4 a, int4 b) {
    asm volatile("st.global.cg.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// store flavours for the record write-back (argv[1] = 0..5)
__device__ __forceinline__ void st32_wb(void* p, int4 a, int4 b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void st32_cs(void* p, int4 a, int4 b) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void st32_hint(void* p, int4 a, int4 b, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z),
                 "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st16x4(void* p, int4 a, int4 b, int4 c, int4 d) {
    int4* q = reinterpret_cast<int4*>(p);
    __stcg(q, a); __stcg(q + 1, b); __stcg(q + 2, c); __stcg(q + 3, d);
}
constexpr size_t kBlock = 4352;  // 256 B keys + 64 x 64 B records
#ifndef PROBE_STAGES
#define PROBE_STAGES 4
#endif
constexpr int kStages = PROBE_STAGES;
constexpr int kStageBytes = 4 * 256 + 32 * 64 + 128;
constexpr int kSmem = 4 * kStages * kStageBytes;

enum { P_KEYS = 1, P_REC_R = 2, P_REC_W = 4, P_POSE = 8, P_ASSOC = 16 };

// which record blob k of particle p touches in frame f (eight distinct records of 64, the same for every particle of
// a frame up to a per-particle rotation -- the real kernel's hits follow the frame's eight visible landmarks)
__device__ __forceinline__ int rec_of(int p, int k, int f, int same) {
    const int rot = same ? 0 : (int)(((unsigned)p * 2654435761u) >> 26);
    return (7 * f + 8 * k + rot) & 63;
}

// one warp = 4 particles x 8 blobs per step; kStages steps in flight
__global__ void __launch_bounds__(128) probe_kernel(unsigned char* pool, const int* slot, double* pose4, int* assoc, int M,
                                                    int what, int frame, int same, unsigned* sink, int flavour) {
    extern __shared__ __align__(128) unsigned char ring_raw[];
    unsigned char (*ring)[kStages][kStageBytes] = reinterpret_cast<unsigned char (*)[kStages][kStageBytes]>(ring_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 4 + warp, tw = gridDim.x * 4;
    const int ngroups = M / 4;
    unsigned acc = 0;
    auto issue = [&](int g, int st) {
        if (g < ngroups) {
            const int p0 = g * 4;
            unsigned char* rb = ring[warp][st];
            if (what & P_KEYS) {
                for (int h = 0; h < 2; ++h) {
                    const int pl = 2 * h + (lane >> 4);
                    const unsigned char* src = pool + (size_t)slot[p0 + pl] * kBlock + 16 * (lane & 15);
                    cpa16(smem_u32(rb + pl * 256 + 16 * (lane & 15)), src);
                }
            }
            if (what & P_REC_R) {
                // cooperative strip fetch as in K2: instruction i moves chunk i*32+lane of the 32-record strip
                const int pl = lane >> 3, k = lane & 7;
                const size_t off = (size_t)slot[p0 + pl] * kBlock + 256 + (size_t)rec_of(p0 + pl, k, frame, same) * 64;
                const unsigned off32 = (unsigned)(off >> 5);
                for (int i = 0; i < 4; ++i) {
                    const int chunk = i * 32 + lane, rec = chunk >> 2, part = chunk & 3;
                    const unsigned so = __shfl_sync(0xffffffffu, off32, rec);
                    cpa16_64(smem_u32(rb + 1024 + 16 * chunk), pool + ((size_t)so << 5) + 16 * part);
                }
            }
            if ((what & P_POSE) && lane < 8) cpa16(smem_u32(rb + 1024 + 2048 + 16 * lane), reinterpret_cast<unsigned char*>(pose4 + 4 * (size_t)p0) + 16 * lane);
        }
        commit();
    };
    int g_issue = gw;
    for (int s = 0; s < kStages - 1; ++s, g_issue += tw) issue(g_issue, s);
    int st = 0;
    for (int g = gw; g < ngroups; g += tw) {
        issue(g_issue, (st + kStages - 1) % kStages);
        g_issue += tw;
        wait_group<kStages - 1>();
        __syncwarp();
        const int p0 = g * 4;
        unsigned char* rb = ring[warp][st];
        const int pl = lane >> 3, k = lane & 7;
        if (what & P_KEYS) acc += reinterpret_cast<unsigned*>(rb + pl * 256)[k * 8];
        int4 r0 = make_int4(lane, g, 0, 1), r1 = r0, r2 = r0, r3 = r0;
        if (what & P_REC_R) {
            const int4* rp = reinterpret_cast<const int4*>(rb + 1024 + 64 * lane);
            r0 = rp[0]; r1 = rp[1]; r2 = rp[2]; r3 = rp[3];
            r0.x += 1;
        }
        if (what & P_REC_W) {
            unsigned char* dst = pool + (size_t)slot[p0 + pl] * kBlock + 256 + (size_t)rec_of(p0 + pl, k, frame, same) * 64;
            if (flavour == 0) { st32(dst, r0, r1); st32(dst + 32, r2, r3); }
            else if (flavour == 1) { st32_wb(dst, r0, r1); st32_wb(dst + 32, r2, r3); }
            else if (flavour == 2) { st32_cs(dst, r0, r1); st32_cs(dst + 32, r2, r3); }
            else if (flavour == 3 || flavour == 4) {
                uint64_t pol;
                if (flavour == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
                st32_hint(dst, r0, r1, pol); st32_hint(dst + 32, r2, r3, pol);
            } else st16x4(dst, r0, r1, r2, r3);
        }
        if (what & P_POSE) {
            acc += reinterpret_cast<unsigned*>(rb + 1024 + 2048)[lane];
            if (k == 0) pose4[4 * (size_t)(p0 + pl) + 3] = (double)acc;
        }
        if (what & P_ASSOC) assoc[(size_t)p0 * 8 + lane] = lane + g;
        __syncwarp();
        st = (st + 1) % kStages;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void copy_kernel(const int4* __restrict__ src, int4* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main(int argc, char** argv) {
    const int M = 1 << 20;
    const int flavour = argc > 1 ? atoi(argv[1]) : 0;
    const int only_perm = argc > 2 ? atoi(argv[2]) : -1;
    printf("stages %d, ", kStages);
    printf("store flavour %d (0 st.cg.v8, 1 st.v8, 2 st.cs.v8, 3 L2 evict_first, 4 L2 evict_last, 5 4 x st.cg.v4)\n", flavour);
    unsigned char* pool;
    int *slot, *assoc;
    double* pose4;
    unsigned* sink;
    CK(cudaMalloc(&pool, (size_t)M * kBlock));
    CK(cudaMemset(pool, 1, (size_t)M * kBlock));
    CK(cudaMalloc(&slot, M * sizeof(int)));
    CK(cudaMalloc(&assoc, (size_t)M * 8 * sizeof(int)));
    CK(cudaMalloc(&pose4, (size_t)M * 32));
    CK(cudaMemset(pose4, 0, (size_t)M * 32));
    CK(cudaMalloc(&sink, 4));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    struct Pat { const char* name; int what; double bytes_pp; };
    const Pat pats[] = {
        {"all", P_KEYS | P_REC_R | P_REC_W | P_POSE | P_ASSOC, 32 + 8 + 256 + 8 * 64 + 8 * 64 + 8 + 32},
        {"keys", P_KEYS, 256},
        {"recs_r", P_REC_R, 8 * 64},
        {"recs_rw", P_REC_R | P_REC_W, 16 * 64},
        {"keys+recs_r", P_KEYS | P_REC_R, 256 + 8 * 64},
    };
    for (int perm = 0; perm < 2; ++perm) {
        if (only_perm >= 0 && perm != only_perm) continue;
        std::vector<int> h(M);
        for (int i = 0; i < M; ++i) h[i] = i;
        if (perm) {  // what copy-on-resample leaves behind: a random permutation of the blocks
            unsigned long long s = 88172645463325252ull;
            for (int i = M - 1; i > 0; --i) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                std::swap(h[i], h[(int)(s % (unsigned long long)(i + 1))]);
            }
        }
        CK(cudaMemcpy(slot, h.data(), M * sizeof(int), cudaMemcpyHostToDevice));
        for (int same = (only_perm >= 0 ? 1 : 0); same < 2; ++same)
            for (const Pat& p : pats)
                for (int ctas = (only_perm >= 0 ? 4 : 2); ctas <= 4; ctas += 2) {  // 8 / 16 warps per SM
                    const int grid = sms * ctas;
                    float best = 1e9f, sum = 0.f;
                    const int reps = 12;
                    for (int it = 0; it < reps + 3; ++it) {
                        CK(cudaEventRecord(e0));
                        probe_kernel<<<grid, 128, kSmem>>>(pool, slot, pose4, assoc, M, p.what, it, same, sink, flavour);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        float ms;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (it >= 3) { best = std::min(best, ms); sum += ms; }
                    }
                    CK(cudaGetLastError());
                    const double gb = p.bytes_pp * M / 1e9;
                    printf("slots=%s records=%s pattern=%-12s warps/SM=%2d  avg %.4f ms  best %.4f ms  %.0f GB/s (algorithmic %.3f GB)\n",
                           perm ? "permuted" : "identity", same ? "same-8" : "rotated", p.name, ctas * 4, sum / reps, best,
                           gb / (sum / reps * 1e-3), gb);
                }
    }
    // the same number of bytes as `all`, contiguous
    {
        const size_t bytes = (size_t)M * 712;  // read 712 + write 712 ~ 1424 B per particle
        const size_t n = bytes / 16;
        int4* src = reinterpret_cast<int4*>(pool);
        int4* dst = reinterpret_cast<int4*>(pool + ((size_t)M * kBlock / 2 & ~(size_t)255));
        float sum = 0.f;
        for (int it = 0; it < 13; ++it) {
            CK(cudaEventRecord(e0));
            copy_kernel<<<sms * 8, 256>>>(src, dst, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 3) sum += ms;
        }
        printf("pattern=stream (copy of %.3f GB read + %.3f GB written)  avg %.4f ms  %.0f GB/s\n", bytes / 1e9, bytes / 1e9,
               sum / 10, 2.0 * bytes / 1e9 / (sum / 10 * 1e-3));
    }
    return 0;
}
