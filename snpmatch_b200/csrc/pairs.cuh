// the two callers next to the matching path that work on the resident panel.
//   k_read_columns  — `g_acc.snps[:, ix]` (simulate.py:15,36-37; genotype_cross.py:97-98; csmatch.py:116-117): whole accession
//                     columns out of the row-major 2-bit panel.  The reference keeps a second, column-chunked HDF5 file for
//                     this access; here one resident copy serves both directions: a column read touches one 32-byte sector
//                     per row (N * 32 bytes of DRAM traffic for any number of columns that share a sector).
//   k_pair_counts   — the per-chromosome agreement counts of pairwiseScore (snpmatch.py:291-297): over the matched marker
//                     pairs of two samples, common[c] = pairs on chromosome c, matches[c] = pairs whose genotype strings
//                     are equal (strings travel as integer ids of the distinct strings of both samples).
#pragma once
#include "common.cuh"

namespace snpm {

constexpr int RC_MAX_COLS = 16;          // columns per launch of k_read_columns

struct ColumnSel {
    int32_t word[RC_MAX_COLS];           // 64-bit word of the row that holds the column
    int32_t bit[RC_MAX_COLS];            // bit inside the two 32-bit planes
    int32_t n;
};

// thread = row; out[c, row] = code of accession sel[c] (3 -> -1).  Consecutive threads write consecutive bytes of every
// output column; the reads of a warp are 32 rows x one sector.
__global__ void __launch_bounds__(256) k_read_columns(const uint64_t *__restrict__ packed, int64_t n_rows, int32_t stride,
                                                      const ColumnSel sel, int8_t *__restrict__ out) {
    for (int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; row < n_rows; row += int64_t(gridDim.x) * blockDim.x) {
        const uint64_t *r = packed + row * stride;
#pragma unroll 4
        for (int c = 0; c < sel.n; ++c) {
            const uint64_t v = __ldg(r + sel.word[c]);
            const uint32_t lo = (uint32_t(v) >> sel.bit[c]) & 1u, hi = (uint32_t(v >> 32) >> sel.bit[c]) & 1u;
            const uint32_t code = lo | (hi << 1);
            out[int64_t(c) * n_rows + row] = code == 3u ? int8_t(-1) : int8_t(code);
        }
    }
}

constexpr int PC_MAX_CHR = 256;          // chromosomes counted in shared memory; more fall back to global atomics

// idx1/idx2: matched pairs (marker indices into the two samples); chrom1: chromosome id of every marker of sample 1
// (ids >= n_chr or < 0 are not counted); gt1/gt2: genotype-string ids.  counts[2 * n_chr] = common | matches.
__global__ void __launch_bounds__(256) k_pair_counts(const int64_t *__restrict__ idx1, const int64_t *__restrict__ idx2, int64_t m,
                                                     const int32_t *__restrict__ chrom1, const int32_t *__restrict__ gt1,
                                                     const int32_t *__restrict__ gt2, int32_t n_chr,
                                                     unsigned long long *__restrict__ counts) {
    __shared__ unsigned int s_common[PC_MAX_CHR], s_match[PC_MAX_CHR];
    const bool in_smem = n_chr <= PC_MAX_CHR;
    if (in_smem) {
        for (int c = threadIdx.x; c < n_chr; c += blockDim.x) {
            s_common[c] = 0u;
            s_match[c] = 0u;
        }
        __syncthreads();
    }
    // a CTA handles at most 2^31 pairs between flushes: the shared counters are 32 bit
    for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < m; k += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = idx1[k], j = idx2[k];
        const int32_t c = chrom1[i];
        if (c < 0 || c >= n_chr) continue;
        const bool same = gt1[i] == gt2[j];
        if (in_smem) {
            atomicAdd(s_common + c, 1u);
            if (same) atomicAdd(s_match + c, 1u);
        } else {
            atomicAdd(counts + c, 1ull);
            if (same) atomicAdd(counts + n_chr + c, 1ull);
        }
    }
    if (in_smem) {
        __syncthreads();
        for (int c = threadIdx.x; c < n_chr; c += blockDim.x) {
            if (s_common[c]) atomicAdd(counts + c, (unsigned long long)s_common[c]);
            if (s_match[c]) atomicAdd(counts + n_chr + c, (unsigned long long)s_match[c]);
        }
    }
}

}  // namespace snpm
