// A2 for called genotypes — popcount scoring of samples whose weights are one-hot (BED inputs and VCFs without PL:
// ParseInputs.get_wei_from_GT, parsers.py:132-139).  With 0/1 weights matchGTsAccs (snpmatch.py:74-89) reduces to
//   score[a] = #rows whose database call equals the sample's call,   ninfo[a] = #rows with a called genotype,
// sums of exact 1.0s, so every summation order gives the reference's fp64 result bit for bit.  No fp64 arithmetic is
// needed: a thread owns one 32-accession word of a row slice, builds per row the match plane
//   M = ~((lo ^ s_lo) | (hi ^ s_hi))          (s_lo/s_hi = the sample's 2-bit code broadcast to all 32 bits)
// and the called plane, and adds them into bit-sliced vertical counters with a Harley-Seal carry-save tree (8 rows per
// step, ~3 LOP3 per row and counter).  The kernel moves 8 bytes per 32 comparisons and needs ~9 integer instructions for
// them: it is bound by the HBM row gather, not by the SM.
#pragma once
#include "common.cuh"

namespace snpm {

constexpr int HC_THREADS = 256;
constexpr int HC_MAX_SEGS = 8;                   // segments per CTA (bounds the staged rows in shared memory: 40 KB)
constexpr int HC_PLANES = 10;                    // counts < 1024 >= SNPM_CHUNK_ROWS

struct HardArgs {
    const uint64_t *packed;
    int32_t stride;
    const int32_t *pair_db;
    const uint8_t *pair_code;     // sample call per matched pair: 0 ref, 1 alt, 2 het; 255 = weights are not one-hot
    const int32_t *seg_off;       // [S+1] chunk mode (as k_score_segments)
    const int32_t *mstart;        // [S+1]
    int32_t S;
    int32_t chunk;
    double *part_score;           // [nseg, a_pad]  (exact integers)
    int32_t *part_ninfo;          // [nseg, a_pad]
    int32_t a_pad;
    int *status;                  // status[3] += rows whose weights were not one-hot
    int32_t wx;                   // words per CTA slice = min(stride, HC_THREADS)
    int32_t spc;                  // segments per CTA = min(HC_THREADS / wx, HC_MAX_SEGS)
};

// full adder on bit planes: (h, l) = a + b + c
#define SNPM_CSA(h, l, a, b, c)                         \
    do {                                                \
        const uint32_t u__ = (a) ^ (b);                 \
        (h) = ((a) & (b)) | (u__ & (c));                \
        (l) = u__ ^ (c);                                \
    } while (0)

struct VCounter {                 // bit-sliced counter of 32 lanes
    uint32_t p[HC_PLANES];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < HC_PLANES; ++k) p[k] = 0u;
    }
    // add eight 1-bit planes (Harley-Seal carry-save tree, then ripple the eights)
    __device__ __forceinline__ void add8(const uint32_t (&m)[8]) {
        uint32_t ta, tb, fa, fb, e;
        SNPM_CSA(ta, p[0], p[0], m[0], m[1]);
        SNPM_CSA(tb, p[0], p[0], m[2], m[3]);
        SNPM_CSA(fa, p[1], p[1], ta, tb);
        SNPM_CSA(ta, p[0], p[0], m[4], m[5]);
        SNPM_CSA(tb, p[0], p[0], m[6], m[7]);
        SNPM_CSA(fb, p[1], p[1], ta, tb);
        SNPM_CSA(e, p[2], p[2], fa, fb);
#pragma unroll
        for (int k = 3; k < HC_PLANES; ++k) {
            const uint32_t c = p[k] & e;
            p[k] ^= e;
            e = c;
        }
    }
    __device__ __forceinline__ int value(int b) const {
        int v = 0;
#pragma unroll
        for (int k = 0; k < HC_PLANES; ++k) v |= int((p[k] >> b) & 1u) << k;
        return v;
    }
};

// grid.x = ceil(segments / segments-per-CTA), grid.y = word slices of HC_THREADS words.  A CTA scores
// HC_THREADS / wx segments side by side (wx = words of its slice): thread -> (segment q, word w), no cross-thread reduction.
template <bool SKIP_HETS>
__global__ void __launch_bounds__(HC_THREADS) k_score_hard(const HardArgs a) {
    const int wx = a.wx, spc = a.spc;
    const int q = threadIdx.x / wx, w = threadIdx.x - q * wx;
    const int seg = blockIdx.x * spc + q;
    const int word = blockIdx.y * wx + w;
    const bool active = q < spc && seg < a.seg_off[a.S];
    const bool has_word = word < a.stride;
    int begin = 0, end = 0;
    if (active) {
        int lo_s = 0, hi_s = a.S;
        while (lo_s < hi_s) {
            const int mid = (lo_s + hi_s + 1) >> 1;
            if (a.seg_off[mid] <= seg) lo_s = mid; else hi_s = mid - 1;
        }
        begin = a.mstart[lo_s] + (seg - a.seg_off[lo_s]) * a.chunk;
        end = min(a.mstart[lo_s + 1], begin + a.chunk);
    }
    const uint64_t *col = a.packed + (has_word ? word : 0);

    // stage the segment's row numbers and sample calls in shared memory (coalesced, once): the gathers below then have
    // no dependent index load in front of them
    extern __shared__ unsigned char hc_smem[];
    const int qs = q < spc ? q : 0;
    int32_t *s_row = reinterpret_cast<int32_t *>(hc_smem) + qs * a.chunk;
    uint8_t *s_code = hc_smem + size_t(spc) * a.chunk * 4 + qs * a.chunk;
    const int n_rows = end - begin;
    for (int r = w; r < n_rows; r += wx) {
        s_row[r] = a.pair_db[begin + r];
        s_code[r] = a.pair_code[begin + r];
    }
    // the threads of one segment share its staged rows; a CTA-wide barrier is safe because no thread has left yet
    __syncthreads();

    VCounter cs, cn;
    cs.clear();
    cn.clear();
    int bad = 0;
    constexpr int HC_DEPTH = 16;                 // independent 8-byte gathers in flight per thread
    for (int r0 = 0; r0 < n_rows; r0 += HC_DEPTH) {
        uint64_t v[HC_DEPTH];
        uint32_t code[HC_DEPTH];
#pragma unroll
        for (int k = 0; k < HC_DEPTH; ++k) {
            const int r = r0 + k;
            const bool ok = r < n_rows;
            v[k] = ok ? __ldg(col + int64_t(s_row[r]) * a.stride) : ~0ull;
            code[k] = ok ? uint32_t(s_code[r]) : 3u;
        }
#pragma unroll
        for (int h = 0; h < HC_DEPTH; h += 8) {
            uint32_t m[8], c[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t lo = uint32_t(v[h + k]), hi = uint32_t(v[h + k] >> 32);
                uint32_t cd = code[h + k];
                if (cd == 255u) { ++bad; cd = 3u; }
                const uint32_t s_lo = 0u - (cd & 1u), s_hi = 0u - (cd >> 1);
                const uint32_t called = SKIP_HETS ? ~hi : ~(lo & hi);   // snpmatch.py:78-79: masked hets are not informative
                m[k] = ~((lo ^ s_lo) | (hi ^ s_hi)) & called;           // database call == sample call
                c[k] = cd == 3u ? 0u : called;                          // padding rows of the last step count nothing
            }
            cs.add8(m);
            cn.add8(c);
        }
    }
    if (bad) atomicAdd(a.status + 3, bad);
    if (!active || !has_word) return;
    double *ps = a.part_score + int64_t(seg) * a.a_pad + int64_t(word) * 32;
    int32_t *pn = a.part_ninfo + int64_t(seg) * a.a_pad + int64_t(word) * 32;
#pragma unroll 4
    for (int b = 0; b < 32; ++b) {
        ps[b] = double(cs.value(b));
        pn[b] = cn.value(b);
    }
}

// sample call of every matched pair from its weights: one-hot (1,0,0) / (0,1,0) / (0,0,1) in (ref, alt, het) order of
// pair_w -> 0 / 1 / 2, anything else 255
__global__ void __launch_bounds__(256) k_pair_codes(const double *__restrict__ pair_w, const int32_t *__restrict__ m_ptr,
                                                    uint8_t *__restrict__ pair_code) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= *m_ptr) return;
    const double4 w = reinterpret_cast<const double4 *>(pair_w)[i];
    uint8_t c = 255;
    if (w.x == 1.0 && w.y == 0.0 && w.z == 0.0) c = 0;
    else if (w.x == 0.0 && w.y == 1.0 && w.z == 0.0) c = 1;
    else if (w.x == 0.0 && w.y == 0.0 && w.z == 1.0) c = 2;
    pair_code[i] = c;
}

}  // namespace snpm
