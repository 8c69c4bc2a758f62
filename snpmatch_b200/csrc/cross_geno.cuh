// windowed parent matching of `snpmatch genotype_cross` (genotype_cross.py:210-241 with
// get_window_genotype_gts :188-199 and getWindowGenotype :21-49).  For every genome window and every sample of a
// multi-sample VCF: among the markers of the window that segregate between the two parents, how many of the sample's calls
// equal parent 1, equal parent 2, or are heterozygous; then the three-way likelihood call (0 = parent 1, 1 = het,
// 2 = parent 2, NA).  The reference runs np.vectorize'd likelihoods per (window, sample) cell; here one CTA owns a window
// and a tile of samples, thread = sample, and the call is fused behind the counts.
#pragma once
#include "common.cuh"
#include "score.cuh"

namespace snpm {

constexpr int GC_THREADS = 128;
constexpr int GC_STAGE = 256;            // pairs staged per step

// getWindowGenotype (genotype_cross.py:21-49).  matched = (parent 1, het, parent 2).  Returns 0/1/2 or -1 (NA).
// *border: the decision hangs on lr_next >= lr_thres within 1e-9 relative (the reference's own rounding would decide).
__device__ __forceinline__ int window_genotype(int c1, int ch, int c2, int total, double lr_thres, int n_marker_thres, int *border) {
    *border = 0;
    if (total < n_marker_thres) return -1;
    if (c1 == 0 && ch == 0 && c2 == 0) return -1;
    double L[3] = {likeli_test(double(total), double(c1)), likeli_test(double(total), double(ch)), likeli_test(double(total), double(c2))};
    double top = nan("");
    int high = -1;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (L[i] == L[i] && !(top <= L[i])) {          // first index of the nan-ignoring minimum (np.nanargmin)
            top = L[i];
            high = i;
        }
    double LR[3];
    int n_one = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        LR[i] = (top <= 0.0) ? nan("") : L[i] / top;     // get_fraction(L, TopHit), snpmatch.py:113-116
        if (LR[i] == 1.0) ++n_one;
    }
    if (n_one > 1) return 1;                             // matching to multiple
    double lr_next = nan("");
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (LR[i] == LR[i] && LR[i] - 1.0 != 0.0 && !(lr_next <= LR[i])) lr_next = LR[i];
    if (lr_next != lr_next) lr_next = lr_thres;
    else if (fabs(lr_next - lr_thres) <= 1e-9 * fabs(lr_thres) && high != 1) *border = 1;
    int geno = -1;
    if (high == 0 && lr_next >= lr_thres) geno = 0;
    else if (high == 2 && lr_next >= lr_thres) geno = 2;
    if (high == 1) geno = 1;
    return geno;
}

// grid (W, ceil(S / GC_THREADS)).  win_start int32[W + 1] cuts the matched pairs (ordered by window) into windows.
// gt int8 [n_vcf, S] row-major (parseGT codes: 0, 1, 2, -1): a warp reads 32 consecutive bytes of a marker's row.
// counts int32 [W, S, 3] = (parent 1, het, parent 2); geno int8 [W, S]; border uint8 [W, S].
__global__ void __launch_bounds__(GC_THREADS) k_gc_window_calls(const int64_t *__restrict__ par_idx, const int64_t *__restrict__ vcf_idx,
                                                                const int32_t *__restrict__ win_start, const int8_t *__restrict__ p1,
                                                                const int8_t *__restrict__ p2, const int8_t *__restrict__ gt, int32_t S,
                                                                double lr_thres, int n_marker_thres, int32_t *__restrict__ counts,
                                                                int8_t *__restrict__ geno, uint8_t *__restrict__ border) {
    __shared__ int64_t s_row[GC_STAGE];
    __shared__ int8_t s_p1[GC_STAGE], s_p2[GC_STAGE];
    const int w = blockIdx.x;
    const int s = blockIdx.y * GC_THREADS + threadIdx.x;
    const int k0 = win_start[w], k1 = win_start[w + 1];
    int c1 = 0, ch = 0, c2 = 0;
    for (int base = k0; base < k1; base += GC_STAGE) {
        const int n = min(GC_STAGE, k1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += GC_THREADS) {
            const int64_t ip = par_idx[base + i];
            s_row[i] = vcf_idx[base + i] * int64_t(S);
            s_p1[i] = p1[ip];
            s_p2[i] = p2[ip];
        }
        __syncthreads();
        if (s < S) {
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                const int g = gt[s_row[i] + s];
                c1 += g == s_p1[i];
                c2 += g == s_p2[i];
                ch += g == 2;
            }
        }
    }
    if (s < S) {
        const int64_t o = int64_t(w) * S + s;
        counts[3 * o + 0] = c1;
        counts[3 * o + 1] = ch;
        counts[3 * o + 2] = c2;
        int b;
        geno[o] = int8_t(window_genotype(c1, ch, c2, k1 - k0, lr_thres, n_marker_thres, &b));
        border[o] = uint8_t(b);
    }
}

}  // namespace snpm
