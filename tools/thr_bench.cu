// Micro-benchmark and self-check of K3b (pk_resample_thresholds) at a given number of scan blocks.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -o tools/thr_bench tools/thr_bench.cu \
//        -Lparakeet_slam_b200 -lparakeet_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../parakeet_slam_b200'
//   tools/thr_bench 8192
//
// Prints the time per launch, and checks the double-double block prefixes against an 80-bit host sum and every
// emitted-output count against its definition #{k : u0 + k r <= prefix}.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "parakeet_b200.h"
int main(int argc, char** argv) {
    long long nb = argc > 1 ? atoll(argv[1]) : 8192;
    long long M = nb * 1024;
    std::vector<double> h(nb);
    srand(1);
    for (auto& v : h) v = 500.0 + (rand() % 1000) * 0.01;
    double *sums, *plan, *prefix; long long* count;
    cudaMalloc(&sums, nb * 8); cudaMalloc(&plan, 8 * 8); cudaMalloc(&prefix, nb * 16); cudaMalloc(&count, (nb + 1) * 8);
    cudaMemcpy(sums, h.data(), nb * 8, cudaMemcpyHostToDevice);
    cudaMemset(plan, 0, 8 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 5; ++i) pk_resample_thresholds(sums, nb, M, 0.37, plan, prefix, count, nullptr);
    cudaEventRecord(a);
    const int R = 50;
    for (int i = 0; i < R; ++i) pk_resample_thresholds(sums, nb, M, 0.37, plan, prefix, count, nullptr);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double hp[8]; cudaMemcpy(hp, plan, 8 * 8, cudaMemcpyDeviceToHost);
    std::vector<long long> hc(nb + 1); cudaMemcpy(hc.data(), count, (nb + 1) * 8, cudaMemcpyDeviceToHost);
    unsigned long long chk = 0; for (auto v : hc) chk = chk * 1315423911ull + (unsigned long long)v;
    printf("nb %lld: %.2f us per launch; total %.6f r %.9g checksum %llx err=%s\n", nb, ms * 1000 / R, hp[0], hp[1], chk, cudaGetErrorString(cudaGetLastError()));
    std::vector<double> hpre(2 * nb); cudaMemcpy(hpre.data(), prefix, nb * 16, cudaMemcpyDeviceToHost);
    long double acc = 0; double worst = 0; long long bad = 0, prev = 0;
    const double r = hp[1], u0 = hp[2];
    for (long long b2 = 0; b2 < nb; ++b2) {
        long double got = (long double)hpre[2 * b2] + (long double)hpre[2 * b2 + 1];
        double e = (double)fabsl(got - acc); if (e > worst) worst = e;
        acc += h[b2];
        // emitted count at the end of block b2: #k in [0,M) with u0 + k r <= acc
        long long want = (long long)floorl((acc - (long double)u0) / (long double)r) + 1; if (want > M) want = M; if (want < 0) want = 0;
        if (b2 == nb - 1) want = M;
        if (hc[b2 + 1] != want) { if (bad < 5) printf("  count[%lld] = %lld want %lld\n", b2 + 1, hc[b2 + 1], want); ++bad; }
        if (hc[b2 + 1] < prev) { printf("  not monotone at %lld\n", b2); ++bad; } prev = hc[b2 + 1];
    }
    printf("  prefix worst abs error vs long double %.3g (total %.3Lg), count mismatches %lld, count[0]=%lld\n", worst, acc, bad, hc[0]);
    return 0;
}
