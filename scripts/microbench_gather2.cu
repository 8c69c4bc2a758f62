// Row-gather microbenchmark, round 2 (measurement tool, not product code).  Questions:
//   (1) does cudaLimitMaxL2FetchGranularity (32 / 64 / 128 B) change what a gather of 288-byte rows costs?  (round 1 measured
//       378 B of DRAM reads per 288-byte row in k_score_grouped: whole 128-byte lines)
//   (2) what is the ceiling in the REAL grouped order (rows ordered by weight-triple id, position order inside a group), not a
//       full shuffle?
//   (3) the kernel's own access shape: per-thread LDGSTS.128 into a shared-memory ring, 4 blocks of 16 rows in flight.
// Prints algorithmic GB/s at 288 B per row.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/microbench_gather2.bin scripts/microbench_gather2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ROW_BYTES = 288;

__global__ void k_fill(uint64_t *p, size_t n) {
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t step = size_t(gridDim.x) * blockDim.x;
    for (; i < n; i += step) p[i] = i * 0x9E3779B97F4A7C15ull;
}

// V1: thread = 8-byte word column, DEPTH independent loads in flight
template <int DEPTH>
__global__ void __launch_bounds__(256) k_v1(const uint64_t *__restrict__ base, int stride_words, const int32_t *__restrict__ rows,
                                            int64_t n_rows, int chunk, uint64_t *sink) {
    const int wx = 36, spc = 7;
    const int q = threadIdx.x / wx, w = threadIdx.x - q * wx;
    if (q >= spc) return;
    const int64_t seg = int64_t(blockIdx.x) * spc + q;
    const int64_t begin = seg * chunk;
    const int64_t end = min(n_rows, begin + chunk);
    uint64_t acc = 0;
    for (int64_t r0 = begin; r0 < end; r0 += DEPTH) {
        uint64_t v[DEPTH];
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) {
            const int64_t r = r0 + k;
            v[k] = r < end ? __ldg(base + int64_t(rows[r]) * stride_words + w) : 0ull;
        }
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) acc ^= v[k];
    }
    if (acc == 0x123456789ull) sink[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// V4: the kernel's shape.  Teams of 36 threads, one segment each; thread pairs copy 16 bytes (two columns) per row into a
// RING-row shared-memory ring, blocks of 16 rows, RING/16 blocks in flight; consumers xor their column.
template <int THREADS, int RING>
__global__ void __launch_bounds__(THREADS) k_v4(const uint64_t *__restrict__ base, int stride_words, const int32_t *__restrict__ rows,
                                                int64_t n_rows, int chunk, uint64_t *sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int WX = 36, SPC = THREADS / WX, INF = RING / 16;
    const int q = threadIdx.x / WX, w = threadIdx.x - q * WX;
    if (q >= SPC) return;
    uint64_t *ring = reinterpret_cast<uint64_t *>(sm) + size_t(q) * RING * WX;
    const int64_t seg = int64_t(blockIdx.x) * SPC + q;
    const int64_t begin = seg * chunk;
    const int64_t end = min(n_rows, begin + chunk);
    const int n_blocks = int((end - begin + 15) / 16);
    const int odd = w & 1;
    const unsigned char *pair_col = reinterpret_cast<const unsigned char *>(base + (w & ~1));
    const uint32_t pair_ring = smem_u32(ring + (w & ~1));
    const uint32_t stride_b = uint32_t(stride_words) * 8u;
    auto issue = [&](int b) {
        if (b < n_blocks) {
            const int64_t r0 = begin + int64_t(b) * 16 + 8 * odd;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t r = r0 + k;
                if (r < end) cp_async16(pair_ring + uint32_t((b * 16 + 8 * odd + k) % RING) * (WX * 8), pair_col + (unsigned long long)(uint32_t(__ldg(rows + r))) * stride_b);
            }
        }
        cp_commit();
    };
#pragma unroll
    for (int b = 0; b < INF; ++b) issue(b);
    uint64_t acc = 0;
    for (int b = 0; b < n_blocks; ++b) {
        cp_wait<INF - 1>();
        __syncwarp();
        const uint64_t *slot = ring + size_t((b * 16) % RING) * WX + w;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc ^= slot[size_t(k) * WX];
        __syncwarp();
        issue(b + INF);
    }
    if (acc == 0x123456789ull) sink[0] = acc;
}

__global__ void __launch_bounds__(256) k_stream(const ulonglong2 *__restrict__ p, size_t n, uint64_t *sink) {
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t step = size_t(gridDim.x) * blockDim.x;
    uint64_t acc = 0;
    for (; i + 3 * step < n; i += 4 * step) {
        ulonglong2 a = __ldg(p + i), b = __ldg(p + i + step), c = __ldg(p + i + 2 * step), d = __ldg(p + i + 3 * step);
        acc ^= a.x ^ a.y ^ b.x ^ b.y ^ c.x ^ c.y ^ d.x ^ d.y;
    }
    if (acc == 0x123456789ull) sink[0] = acc;
}

template <typename F>
static float time_it(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char **argv) {
    const int64_t N = 10700000;
    const int S = 64, M = 45000;
    const int only_gran = argc > 1 ? atoi(argv[1]) : -1;       // one granularity only (for an ncu pass)
    const int64_t n_rows = int64_t(S) * M;
    std::vector<int32_t> h_sorted(n_rows), h_shuf, h_grp(n_rows);
    std::mt19937_64 rng(12345);
    // group sizes of a PL sample: a few big groups, a long tail (Zipf-like over 740 ids)
    std::vector<double> cdf(740);
    double tot = 0;
    for (int g = 0; g < 740; ++g) { tot += 1.0 / (1.0 + g * 0.15); cdf[g] = tot; }
    for (int s = 0; s < S; ++s) {
        std::vector<int32_t> r(M);
        for (int i = 0; i < M; ++i) r[i] = int32_t(int64_t(i) * (N / M) + int64_t(rng() % uint64_t(N / M)));
        std::sort(r.begin(), r.end());
        std::copy(r.begin(), r.end(), h_sorted.begin() + int64_t(s) * M);
        std::vector<std::pair<int, int32_t>> kv(M);
        for (int i = 0; i < M; ++i) {
            const double u = double(rng() >> 11) / 9007199254740992.0 * tot;
            kv[i] = {int(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin()), r[i]};
        }
        std::sort(kv.begin(), kv.end());
        for (int i = 0; i < M; ++i) h_grp[int64_t(s) * M + i] = kv[i].second;
    }
    h_shuf = h_sorted;
    for (int s = 0; s < S; ++s) std::shuffle(h_shuf.begin() + int64_t(s) * M, h_shuf.begin() + int64_t(s + 1) * M, rng);
    int32_t *d_rows[3];
    const char *oname[3] = {"sorted", "grouped", "shuffled"};
    const std::vector<int32_t> *h[3] = {&h_sorted, &h_grp, &h_shuf};
    for (int o = 0; o < 3; ++o) {
        CK(cudaMalloc(&d_rows[o], n_rows * 4));
        CK(cudaMemcpy(d_rows[o], h[o]->data(), n_rows * 4, cudaMemcpyHostToDevice));
    }
    uint64_t *sink;
    CK(cudaMalloc(&sink, 8));
    const double alg = double(n_rows) * ROW_BYTES;
    const int stride = 288, sw = stride / 8;
    uint64_t *base;
    const size_t bytes = size_t(N) * stride;
    CK(cudaMalloc(&base, bytes));
    k_fill<<<148 * 8, 256>>>(base, bytes / 8);
    CK(cudaDeviceSynchronize());
    for (int gran : {0, 32, 64, 128}) {
        if (only_gran >= 0 && gran != only_gran) continue;
        if (gran) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(gran));
            if (e != cudaSuccess) { printf("set granularity %d: %s\n", gran, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        }
        size_t got = 0;
        CK(cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity));
        printf("== L2 fetch granularity requested %d, reported %zu\n", gran, got);
        {
            const size_t n16 = size_t(alg / 16);
            float ms = time_it([&] { k_stream<<<148 * 16, 256>>>(reinterpret_cast<const ulonglong2 *>(base), n16, sink); });
            printf("stream read of %.1f MB: %.3f ms  %.0f GB/s\n", alg / 1e6, ms, alg / ms / 1e6);
        }
        for (int o = 0; o < 3; ++o) {
            const int chunk = 320;
            const int64_t nseg = (n_rows + chunk - 1) / chunk;
            float ms = time_it([&] { k_v1<32><<<unsigned((nseg + 6) / 7), 256>>>(base, sw, d_rows[o], n_rows, chunk, sink); });
            printf("gran %d %s V1 ldg64 depth32 chunk %d: %.3f ms  %.0f GB/s\n", gran, oname[o], chunk, ms, alg / ms / 1e6);
            {
                constexpr int T = 128, R = 64;
                const int smem = (T / 36) * R * 288;
                CK(cudaFuncSetAttribute(k_v4<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                ms = time_it([&] { k_v4<T, R><<<unsigned((nseg + 2) / 3), T, smem>>>(base, sw, d_rows[o], n_rows, chunk, sink); });
                printf("gran %d %s V4 ldgsts ring64 128thr chunk %d: %.3f ms  %.0f GB/s\n", gran, oname[o], chunk, ms, alg / ms / 1e6);
            }
            {
                constexpr int T = 288, R = 64;
                const int smem = (T / 36) * R * 288;
                CK(cudaFuncSetAttribute(k_v4<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                ms = time_it([&] { k_v4<T, R><<<unsigned((nseg + 7) / 8), T, smem>>>(base, sw, d_rows[o], n_rows, chunk, sink); });
                printf("gran %d %s V4 ldgsts ring64 288thr chunk %d: %.3f ms  %.0f GB/s\n", gran, oname[o], chunk, ms, alg / ms / 1e6);
            }
            {
                constexpr int T = 288, R = 32;
                const int smem = (T / 36) * R * 288;
                CK(cudaFuncSetAttribute(k_v4<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                ms = time_it([&] { k_v4<T, R><<<unsigned((nseg + 7) / 8), T, smem>>>(base, sw, d_rows[o], n_rows, chunk, sink); });
                printf("gran %d %s V4 ldgsts ring32 288thr chunk %d: %.3f ms  %.0f GB/s\n", gran, oname[o], chunk, ms, alg / ms / 1e6);
            }
        }
    }
    return 0;
}
