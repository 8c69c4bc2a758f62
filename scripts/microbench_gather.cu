// Row-gather microbenchmark (measurement tool, not product code): how fast can one B200 fetch random packed-panel rows
// (the access pattern of single-sample scoring: 64 samples x 45 000 sorted random rows out of 10.7 M) with
//   V1  8-byte LDG per thread, thread = word column, 16 independent loads in flight (the k_score_hard pattern)
//   V2  16-byte LDG, a warp fetches whole rows (18-20 lanes x 16 B), D rows in flight per warp
//   V3  one 1-D TMA bulk copy per row into a shared-memory ring (64-row tiles), consumers read the tile back
// for row strides of 288 B (32-byte aligned rows), 320 B (64-byte aligned) and 384 B (128-byte aligned).
// Prints algorithmic GB/s counted at 288 B per row for every variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/microbench_gather scripts/microbench_gather.cu
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

// V1: thread = 8-byte word column; CTA of 256 threads = 7 row slots x 36 words; chunk of rows per CTA
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

// V2: warp fetches whole rows with 16-byte loads (lanes 0..17 for 288 B), DEPTH rows in flight
template <int DEPTH>
__global__ void __launch_bounds__(256) k_v2(const uint64_t *__restrict__ base, int stride_words, const int32_t *__restrict__ rows,
                                            int64_t n_rows, int chunk, uint64_t *sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t seg = int64_t(blockIdx.x) * 8 + warp;
    const int64_t begin = seg * chunk;
    const int64_t end = min(n_rows, begin + chunk);
    uint64_t acc = 0;
    const bool on = lane < ROW_BYTES / 16;
    for (int64_t r0 = begin; r0 < end; r0 += DEPTH) {
        ulonglong2 v[DEPTH];
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) {
            const int64_t r = r0 + k;
            v[k] = make_ulonglong2(0, 0);
            if (on && r < end) v[k] = __ldg(reinterpret_cast<const ulonglong2 *>(base + int64_t(rows[r]) * stride_words) + lane);
        }
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) acc ^= v[k].x ^ v[k].y;
    }
    if (acc == 0x123456789ull) sink[0] = acc;
}

// V3: TMA bulk ring
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int STAGES, int TILE>
__global__ void __launch_bounds__(160) k_v3(const uint64_t *__restrict__ base, int stride_words, const int32_t *__restrict__ rows,
                                            int64_t n_rows, int chunk, uint64_t *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + STAGES;
    unsigned char *tiles = smem + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t begin = int64_t(blockIdx.x) * chunk;
    const int64_t end = min(n_rows, begin + chunk);
    const int n_tiles = int((end - begin + TILE - 1) / TILE);
    if (warp == 4) {
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t % STAGES, it = t / STAGES;
            if (it > 0) mbar_wait(empty + s, (it - 1) & 1);
            const int64_t r0 = begin + int64_t(t) * TILE;
            const int cnt = int(min(int64_t(TILE), end - r0));
            if (lane == 0) mbar_expect_tx(full + s, uint32_t(cnt) * ROW_BYTES);
            __syncwarp();
            for (int k = lane; k < cnt; k += 32)
                bulk_g2s(tiles + size_t(s) * TILE * ROW_BYTES + size_t(k) * ROW_BYTES, base + int64_t(rows[r0 + k]) * stride_words,
                         ROW_BYTES, full + s);
        }
    } else {
        uint64_t acc = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t % STAGES, it = t / STAGES;
            mbar_wait(full + s, it & 1);
            const uint64_t *tile = reinterpret_cast<const uint64_t *>(tiles + size_t(s) * TILE * ROW_BYTES);
            const int64_t r0 = begin + int64_t(t) * TILE;
            const int cnt = int(min(int64_t(TILE), end - r0));
            for (int i = threadIdx.x; i < cnt * (ROW_BYTES / 8); i += 128) acc ^= tile[i];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        if (acc == 0x123456789ull) sink[0] = acc;
    }
}

// sequential streaming read of the same byte volume (upper reference)
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
    const int S = argc > 1 ? atoi(argv[1]) : 64, M = 45000;
    const int64_t n_rows = int64_t(S) * M;
    std::vector<int32_t> h_rows(n_rows);
    std::mt19937_64 rng(12345);
    for (int s = 0; s < S; ++s) {
        // sorted distinct random rows
        std::vector<int32_t> r(M);
        for (int i = 0; i < M; ++i) r[i] = int32_t(int64_t(i) * (N / M) + int64_t(rng() % uint64_t(N / M)));
        std::sort(r.begin(), r.end());
        std::copy(r.begin(), r.end(), h_rows.begin() + int64_t(s) * M);
    }
    // a shuffled version (the order a weight-grouped kernel would gather in)
    std::vector<int32_t> h_shuf(h_rows);
    for (int s = 0; s < S; ++s) std::shuffle(h_shuf.begin() + int64_t(s) * M, h_shuf.begin() + int64_t(s + 1) * M, rng);
    int32_t *d_rows, *d_shuf;
    CK(cudaMalloc(&d_rows, n_rows * 4));
    CK(cudaMalloc(&d_shuf, n_rows * 4));
    CK(cudaMemcpy(d_rows, h_rows.data(), n_rows * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_shuf, h_shuf.data(), n_rows * 4, cudaMemcpyHostToDevice));
    uint64_t *sink;
    CK(cudaMalloc(&sink, 8));
    const double alg_bytes = double(n_rows) * ROW_BYTES;
    printf("rows gathered per launch: %lld (%.1f MB at 288 B)\n", (long long)n_rows, alg_bytes / 1e6);
    const int strides[3] = {288, 320, 384};
    for (int si = 0; si < 3; ++si) {
        const int stride = strides[si];
        uint64_t *base;
        const size_t bytes = size_t(N) * stride;
        CK(cudaMalloc(&base, bytes));
        k_fill<<<148 * 8, 256>>>(base, bytes / 8);
        CK(cudaDeviceSynchronize());
        const int sw = stride / 8;
        if (si == 0) {
            const size_t n16 = alg_bytes / 16;
            float ms = time_it([&] { k_stream<<<148 * 16, 256>>>(reinterpret_cast<const ulonglong2 *>(base), n16, sink); });
            printf("stream read of %.1f MB: %.3f ms  %.0f GB/s\n", alg_bytes / 1e6, ms, alg_bytes / ms / 1e6);
        }
        for (int order = 0; order < 2; ++order) {
            const int32_t *rows = order ? d_shuf : d_rows;
            const char *oname = order ? "shuffled" : "sorted";
            for (int chunk : {1000}) {
                const int64_t nseg = (n_rows + chunk - 1) / chunk;
                float ms = time_it([&] { k_v1<16><<<unsigned((nseg + 6) / 7), 256>>>(base, sw, rows, n_rows, chunk, sink); });
                printf("stride %d %s V1 ldg64 depth16 chunk %d: %.3f ms  %.0f GB/s\n", stride, oname, chunk, ms, alg_bytes / ms / 1e6);
                ms = time_it([&] { k_v1<32><<<unsigned((nseg + 6) / 7), 256>>>(base, sw, rows, n_rows, chunk, sink); });
                printf("stride %d %s V1 ldg64 depth32 chunk %d: %.3f ms  %.0f GB/s\n", stride, oname, chunk, ms, alg_bytes / ms / 1e6);
            }
            for (int chunk : {250, 1000}) {
                const int64_t nseg = (n_rows + chunk - 1) / chunk;
                float ms = time_it([&] { k_v2<8><<<unsigned((nseg + 7) / 8), 256>>>(base, sw, rows, n_rows, chunk, sink); });
                printf("stride %d %s V2 ldg128 warp/row depth8 chunk %d: %.3f ms  %.0f GB/s\n", stride, oname, chunk, ms, alg_bytes / ms / 1e6);
                ms = time_it([&] { k_v2<16><<<unsigned((nseg + 7) / 8), 256>>>(base, sw, rows, n_rows, chunk, sink); });
                printf("stride %d %s V2 ldg128 warp/row depth16 chunk %d: %.3f ms  %.0f GB/s\n", stride, oname, chunk, ms, alg_bytes / ms / 1e6);
            }
            for (int chunk : {1000, 4000}) {
                const int64_t nseg = (n_rows + chunk - 1) / chunk;
                {
                    constexpr int ST = 4, TL = 64;
                    const int smem = 128 + ST * TL * ROW_BYTES;
                    CK(cudaFuncSetAttribute(k_v3<ST, TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                    float ms = time_it([&] { k_v3<ST, TL><<<unsigned(nseg), 160, smem>>>(base, sw, rows, n_rows, chunk, sink); });
                    printf("stride %d %s V3 tma ring 4x64 rows chunk %d: %.3f ms  %.0f GB/s\n", stride, oname, chunk, ms, alg_bytes / ms / 1e6);
                }
                {
                    constexpr int ST = 3, TL = 32;
                    const int smem = 128 + ST * TL * ROW_BYTES;
                    CK(cudaFuncSetAttribute(k_v3<ST, TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                    float ms = time_it([&] { k_v3<ST, TL><<<unsigned(nseg), 160, smem>>>(base, sw, rows, n_rows, chunk, sink); });
                    printf("stride %d %s V3 tma ring 3x32 rows chunk %d: %.3f ms  %.0f GB/s\n", stride, oname, chunk, ms, alg_bytes / ms / 1e6);
                }
            }
        }
        CK(cudaFree(base));
    }
    return 0;
}
