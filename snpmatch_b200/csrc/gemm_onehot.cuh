// A9 — batched scoring of many called-genotype samples on a SHARED marker panel as a one-hot int8 GEMM on the 5th-gen
// tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).  North-star item (d); no reference symbol (the reference
// runs one process per sample, README.md:9).
//
//   score[s, a] = sum_k sum_c [code_s[k] == c] * [db[k, a] == c]          c in {ref, alt, het}
//   ninfo[s, a] = sum_k [code_s[k] called] * [db[k, a] called]
//
// Both are one GEMM  C[2*64 rows, 128 accessions] += A_op[128, 4K] * B_op[128, 4K]^T  per CTA tile: every panel row k
// contributes four int8 K-slots (slot c = one-hot of the code, slot 3 unused); the first 64 rows of a tile are the score
// operands of 64 samples (one-hot of the sample's call), the last 64 their ninfo operands (1 in every class slot when the
// sample has the marker).  Neither operand ever exists in HBM: four loader warps expand the 2-bit panel words and the
// samples' code bytes straight into shared memory in the canonical no-swizzle K-major UMMA layout (core matrix = 8 rows x
// 16 bytes, contiguous), fence them into the async proxy, and one elected thread of the MMA warp issues four
// tcgen05.mma (M=128, N=128, K=32) per 32-row block.  int32 accumulation is exact.
#pragma once
#include "common.cuh"
#include "score.cuh"

namespace snpm {

constexpr int OG_STAGES = 4;
constexpr int OG_BM = 128, OG_BN = 128;
constexpr int OG_ROWS = 32;                     // panel rows per k-block = 128 K-bytes = 4 MMAs of K=32
constexpr int OG_TILE_BYTES = 128 * 128;        // one operand tile
constexpr int OG_THREADS = 160;                 // warps 0-3: loaders + epilogue, warp 4: TMEM owner + MMA issuer

struct OneHotGemmArgs {
    const uint64_t *packed;      // panel
    int32_t stride;
    const int32_t *rows;         // [Kpad] local panel rows of the shared markers, -1 = padding
    const uint8_t *codes;        // [S, Kpad] sample calls: 0 ref, 1 alt, 2 het, 3 absent
    int32_t S, Kpad, n_acc;
    int32_t skip_hets;
    int32_t *out_score;          // [S, ld_out]
    int32_t *out_ninfo;          // [S, ld_out]
    int32_t ld_out;              // accessions rounded up to a multiple of OG_BN
};

__device__ __forceinline__ uint64_t og_smem_desc(uint32_t saddr) {
    // K-major, no swizzle: 16-byte rows of a core matrix are contiguous, 8-row groups are 128 B apart (SBO), the two
    // 16-byte K chunks of one MMA are 2048 B apart (LBO).  Bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1
    return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(2048u >> 4) << 16) | (uint64_t(128u >> 4) << 32) | (uint64_t(1) << 46);
}

__global__ void __launch_bounds__(OG_THREADS, 1) k_onehot_gemm(const OneHotGemmArgs a) {
    extern __shared__ __align__(1024) unsigned char og_smem[];
    __shared__ uint64_t full[OG_STAGES], empty[OG_STAGES], acc_ready;
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x, n_blk = blockIdx.y;
    const int n_kb = a.Kpad / OG_ROWS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < OG_STAGES; ++s) {
            mbar_init(smem_u32(&full[s]), 4u);          // one arrive per loader warp
            mbar_init(smem_u32(&empty[s]), 1u);         // tcgen05.commit
        }
        mbar_init(smem_u32(&acc_ready), 1u);
        mbar_fence_init();
    }
    if (warp == 4) {                                    // TMEM: 128 columns of 32-bit accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp < 4) {
        // ---- loaders: expand both operand tiles of k-block kb into stage kb % STAGES ------------------------------
        const int t = threadIdx.x;                       // 0..127 = operand row (m or n)
        const int sample = m_blk * 64 + (t & 63);
        const bool ninfo_row = t >= 64;
        const uint8_t *crow = a.codes + int64_t(min(sample, a.S - 1)) * a.Kpad;
        const bool sample_ok = sample < a.S;
        const int acc = n_blk * OG_BN + t;
        const bool acc_ok = acc < a.stride * 32;          // columns beyond the padded row produce zeros
        const uint64_t *pcol = a.packed + (acc_ok ? (acc >> 5) : 0);
        const int bit = acc & 31;
        const uint32_t row_off = uint32_t(t >> 3) * 128u + uint32_t(t & 7) * 16u;   // inside a K-chunk slab of 16 row groups
        for (int kb = 0; kb < n_kb; ++kb) {
            const int st = kb % OG_STAGES;
            mbar_wait(smem_u32(&empty[st]), (uint32_t(kb / OG_STAGES) & 1u) ^ 1u);
            unsigned char *sa = og_smem + size_t(st) * 2 * OG_TILE_BYTES;
            unsigned char *sb = sa + OG_TILE_BYTES;
#pragma unroll
            for (int c = 0; c < 8; ++c) {                // 16-byte K chunk c = panel rows 4c..4c+3 of the block
                const int k0 = kb * OG_ROWS + 4 * c;
                // A: the sample's four calls
                uint32_t cw = sample_ok ? *reinterpret_cast<const uint32_t *>(crow + k0) : 0x03030303u;
                uint4 va;
                uint32_t *pa = reinterpret_cast<uint32_t *>(&va);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t code = (cw >> (8 * q)) & 0xFFu;
                    pa[q] = code < 3u ? (ninfo_row ? 0x00010101u : (1u << (8 * code))) : 0u;
                }
                *reinterpret_cast<uint4 *>(sa + c * 2048 + row_off) = va;
                // B: the panel's four calls for this accession
                uint4 vb;
                uint32_t *pb = reinterpret_cast<uint32_t *>(&vb);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int32_t r = a.rows[k0 + q];
                    uint32_t w = 0u;
                    if (r >= 0 && acc_ok) {
                        const uint64_t v = __ldg(pcol + int64_t(r) * a.stride);
                        const uint32_t code = (uint32_t(v) >> bit & 1u) | ((uint32_t(v >> 32) >> bit & 1u) << 1);
                        if (code < 3u && !(a.skip_hets && code == 2u)) w = 1u << (8 * code);
                    }
                    pb[q] = w;
                }
                *reinterpret_cast<uint4 *>(sb + c * 2048 + row_off) = vb;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full[st]));
        }
        // ---- epilogue: TMEM -> registers -> global; warp w reads TMEM lanes 32w..32w+31 = tile rows ------------------
        mbar_wait(smem_u32(&acc_ready), 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = warp * 32 + lane;                // tile row: < 64 score of sample, >= 64 ninfo
        const int s_out = m_blk * 64 + (row & 63);
        int32_t *dst = (row < 64 ? a.out_score : a.out_ninfo) + int64_t(s_out) * a.ld_out + n_blk * OG_BN;
#pragma unroll 1
        for (int col = 0; col < OG_BN; col += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_acc + (uint32_t(warp * 32) << 16) + uint32_t(col);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (s_out < a.S) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<int4 *>(dst + col + j) = make_int4(int(r[j]), int(r[j + 1]), int(r[j + 2]), int(r[j + 3]));
            }
        }
    } else {
        // ---- MMA issuer: one elected thread -----------------------------------------------------------------------------
        // instruction descriptor (kind::i8): D = S32 (bits 4-5 = 2), A/B = unsigned 8-bit, both K-major, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (2u << 4) | (uint32_t(OG_BN >> 3) << 17) | (uint32_t(OG_BM >> 4) << 24);
        for (int kb = 0; kb < n_kb; ++kb) {
            const int st = kb % OG_STAGES;
            mbar_wait(smem_u32(&full[st]), uint32_t(kb / OG_STAGES) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sa = smem_u32(og_smem + size_t(st) * 2 * OG_TILE_BYTES), sb = sa + OG_TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {         // K = 32 bytes = K-chunks 2kk, 2kk+1
                    const uint64_t da = og_smem_desc(sa + kk * 4096), db = og_smem_desc(sb + kk * 4096);
                    const uint32_t accumulate = (kb | kk) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem_acc), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
                }
                // the stage is free once these MMAs have read it; the accumulator is complete after the last block
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
                if (kb == n_kb - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&acc_ready)) : "memory");
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_acc) : "memory");
}

// int32 GEMM outputs -> the f64 reduce rows the likelihood epilogue reads (score | ninfo | markers | 0)
__global__ void __launch_bounds__(256) k_onehot_totals(const int32_t *__restrict__ out_score, const int32_t *__restrict__ out_ninfo,
                                                       int32_t ld_out, int32_t n_acc, int32_t k_markers, double *__restrict__ red) {
    const int s = blockIdx.y;
    const int acc = blockIdx.x * blockDim.x + threadIdx.x;
    double *row = red + int64_t(s) * (2 * int64_t(n_acc) + 2);
    if (acc < n_acc) {
        row[acc] = double(out_score[int64_t(s) * ld_out + acc]);
        row[n_acc + acc] = double(out_ninfo[int64_t(s) * ld_out + acc]);
    }
    if (acc == 0) {
        row[2 * n_acc] = double(k_markers);
        row[2 * n_acc + 1] = 0.0;
    }
}

}  // namespace snpm
