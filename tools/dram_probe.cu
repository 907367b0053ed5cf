// Micro-probe: DRAM bytes actually moved by the access patterns of the fused measurement kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dram_probe tools/dram_probe.cu
// Run under: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr size_t kBlock = 4352;  // 256 B keys + 64 x 64 B records
constexpr long long kM = 1 << 20;

// (a) one 256-byte bulk copy per particle (keys)
__global__ void k_tma_keys(const unsigned char* base, long long M, unsigned* sink) {
    __shared__ __align__(128) unsigned char buf[8][256];
    __shared__ uint64_t bar;
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    unsigned ph = 0, acc = 0;
    for (long long p0 = (long long)blockIdx.x * 8; p0 < M; p0 += (long long)gridDim.x * 8) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(8 * 256));
        __syncwarp();
        if (lane < 8)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(buf[lane])),
                         "l"(base + (size_t)(p0 + lane) * kBlock), "r"(256), "r"(smem_u32(&bar))
                         : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(ph) : "memory");
        } while (!ok);
        ph ^= 1;
        acc += reinterpret_cast<unsigned*>(buf[lane & 7])[lane];
        __syncwarp();
    }
    if (acc == 0x12345678u) *sink = acc;
}

__device__ __forceinline__ int rec_of(long long p, int k) { return (int)(((unsigned)(p * 2654435761u) >> 7) + 8 * k) & 63; }

// (b) eight 64-byte records per particle with cp.async 16 B (lane = (particle, blob), 4 copies per lane)
__global__ void k_ldgsts_rec(const unsigned char* base, long long M, unsigned* sink, int fixed) {
    __shared__ __align__(128) unsigned char buf[32][64];
    const int lane = threadIdx.x;
    unsigned acc = 0;
    for (long long p0 = (long long)blockIdx.x * 4; p0 < M; p0 += (long long)gridDim.x * 4) {
        const long long p = p0 + (lane >> 3);
        const int k = lane & 7;
        const int j = fixed ? 8 * k : rec_of(p, k);
        const unsigned char* src = base + (size_t)p * kBlock + 256 + (size_t)j * 64;
        for (int q = 0; q < 4; ++q)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf[lane]) + 16 * q), "l"(src + 16 * q) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        acc += reinterpret_cast<unsigned*>(buf[lane])[lane & 15];
        __syncwarp();
    }
    if (acc == 0x12345678u) *sink = acc;
}

// (c) the same records with ld.global.cg.v4 (two lanes cover one sector in the same instruction)
__global__ void k_ldg_rec(const unsigned char* base, long long M, unsigned* sink, int write_back) {
    const int lane = threadIdx.x;
    unsigned acc = 0;
    for (long long p0 = (long long)blockIdx.x * 4; p0 < M; p0 += (long long)gridDim.x * 4) {
        const long long p = p0 + (lane >> 3);
        const int k = lane & 7;
        const int j = rec_of(p, k);
        unsigned char* src = const_cast<unsigned char*>(base) + (size_t)p * kBlock + 256 + (size_t)j * 64;
        int4 v[4];
        for (int q = 0; q < 4; ++q) v[q] = __ldcg(reinterpret_cast<const int4*>(src) + q);
        for (int q = 0; q < 4; ++q) acc += v[q].x + v[q].w;
        if (write_back) {
            v[0].x += 1;
            asm volatile("st.global.cg.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(src), "r"(v[0].x), "r"(v[0].y), "r"(v[0].z),
                         "r"(v[0].w), "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w) : "memory");
            asm volatile("st.global.cg.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(src + 32), "r"(v[2].x), "r"(v[2].y), "r"(v[2].z),
                         "r"(v[2].w), "r"(v[3].x), "r"(v[3].y), "r"(v[3].z), "r"(v[3].w) : "memory");
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// (d) keys with plain coalesced 16-byte loads (lane reads 16 B; 16 lanes per particle)
__global__ void k_ldg_keys(const unsigned char* base, long long M, unsigned* sink) {
    const int lane = threadIdx.x;
    unsigned acc = 0;
    for (long long p0 = (long long)blockIdx.x * 2; p0 < M; p0 += (long long)gridDim.x * 2) {
        const long long p = p0 + (lane >> 4);
        const int4 v = __ldcg(reinterpret_cast<const int4*>(base + (size_t)p * kBlock) + (lane & 15));
        acc += v.x + v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}


// (e) record reads with explicit L2 prefetch-size / policy variants
template <int MODE>
__global__ void k_rec_var(const unsigned char* base, long long M, unsigned* sink) {
    __shared__ __align__(128) unsigned char buf[32][64];
    __shared__ uint64_t bar;
    const int lane = threadIdx.x;
    unsigned acc = 0, ph = 0;
    if (MODE == 4 || MODE == 6) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;");
        }
        __syncwarp();
    }
    unsigned long long pol = 0;
    if (MODE == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (long long p0 = (long long)blockIdx.x * 4; p0 < M; p0 += (long long)gridDim.x * 4) {
        const long long p = p0 + (lane >> 3);
        const int k = lane & 7;
        const int j = rec_of(p, k);
        const unsigned char* src = base + (size_t)p * kBlock + 256 + (size_t)j * 64;
        int4 v[4] = {};
        if (MODE == 0) {  // ld.global.cg.L2::64B
            for (int q = 0; q < 4; ++q)
                asm volatile("ld.global.cg.L2::64B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v[q].x), "=r"(v[q].y), "=r"(v[q].z), "=r"(v[q].w) : "l"(src + 16 * q));
        } else if (MODE == 1) {  // ld.global.nc
            for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const int4*>(src) + q);
        } else if (MODE == 2) {  // evict_first cache hint
            for (int q = 0; q < 4; ++q)
                asm volatile("ld.global.cg.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v[q].x), "=r"(v[q].y), "=r"(v[q].z), "=r"(v[q].w) : "l"(src + 16 * q), "l"(pol));
        } else if (MODE == 3) {  // only the first 32-byte sector of each record
            for (int q = 0; q < 2; ++q) v[q] = __ldcg(reinterpret_cast<const int4*>(src) + q);
        } else if (MODE == 4) {  // one 64-byte bulk copy per record (each lane its own)
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32 * 64));
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf[lane])),
                         "l"(src), "r"(64), "r"(smem_u32(&bar)) : "memory");
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"(ph) : "memory");
            } while (!ok);
            ph ^= 1;
            v[0] = *reinterpret_cast<int4*>(buf[lane]);
            __syncwarp();
        } else if (MODE == 5) {  // 32-byte loads (ld.global.v8 = LDG.256?)
            for (int q = 0; q < 2; ++q)
                asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[2*q].x), "=r"(v[2*q].y), "=r"(v[2*q].z), "=r"(v[2*q].w),
                             "=r"(v[2*q+1].x), "=r"(v[2*q+1].y), "=r"(v[2*q+1].z), "=r"(v[2*q+1].w) : "l"(src + 32 * q));
        } else if (MODE == 6) {  // records on 128-byte strides (one record per line): is the line still fetched whole?
            const unsigned char* s2 = base + (size_t)p * kBlock + 256 + (size_t)(j >> 1) * 128;
            for (int q = 0; q < 4; ++q) v[q] = __ldcg(reinterpret_cast<const int4*>(s2) + q);
        }
        for (int q = 0; q < 4; ++q) acc += v[q].x + v[q].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    unsigned char* base;
    unsigned* sink;
    cudaMalloc(&base, kM * kBlock);
    cudaMalloc(&sink, 4);
    cudaMemset(base, 1, kM * kBlock);
    const int grid = 148 * 16;
    for (int rep = 0; rep < 2; ++rep) {
        k_tma_keys<<<grid, 32>>>(base, kM, sink);
        k_ldgsts_rec<<<grid, 32>>>(base, kM, sink, 0);
        k_ldgsts_rec<<<grid, 32>>>(base, kM, sink, 1);
        k_ldg_rec<<<grid, 32>>>(base, kM, sink, 0);
        k_ldg_rec<<<grid, 32>>>(base, kM, sink, 1);
        k_ldg_keys<<<grid, 32>>>(base, kM, sink);
        k_rec_var<0><<<grid, 32>>>(base, kM, sink);
        k_rec_var<1><<<grid, 32>>>(base, kM, sink);
        k_rec_var<2><<<grid, 32>>>(base, kM, sink);
        k_rec_var<3><<<grid, 32>>>(base, kM, sink);
        k_rec_var<4><<<grid, 32>>>(base, kM, sink);
        k_rec_var<5><<<grid, 32>>>(base, kM, sink);
        k_rec_var<6><<<grid, 32>>>(base, kM, sink);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    printf("expected per launch: keys %.3f GB, records %.3f GB\n", kM * 256 / 1e9, kM * 512 / 1e9);
    return 0;
}
