// A2/A3/A4 — per-accession scoring of matched rows, chunk combine, likelihood epilogue.
//
// k_score_segments is the hot kernel.  It restates matchGTsAccs (snpmatch.py:74-89) for a list of
// row segments — the 1000-pair chunks of Genotyper.genotyper (snpmatch.py:218-225) or the windows of
// CrossIdentifier.window_genotyper (csmatch.py:80-95) — in the reference's floating-point order:
// per accession and class a plain left-to-right sum over the segment's rows, classes combined as
// ((0 + ref) + het) + alt, so fp64 scores are bit-identical to NumPy's.
//
// Mapping: lane = accession inside a 32-accession word, a warp owns SC_WPW consecutive words, a CTA
// owns a contiguous word slice of every row of ONE segment.  Row slices (gathered rows of the 2-bit
// panel) and the segment's weights travel HBM -> shared memory as 1-D TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx) through an SC_STAGES-deep ring, issued by warp 0 while all
// warps reduce the previous tile from shared memory with broadcast 128-bit loads.
#pragma once
#include "common.cuh"

namespace snpm {

constexpr int SC_TILE_ROWS = 64;
constexpr int SC_STAGES = 3;
constexpr int SC_WPW = 4;          // words (of 32 accessions) per warp
constexpr int SC_MAX_WARPS = 12;

// ---- PTX helpers (sm_100a) ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}

struct ScoreArgs {
    const uint64_t *packed;     // [n_rows, stride]
    int32_t stride;
    const int32_t *pair_db;     // matched local rows
    const double *pair_w;       // [m, 4] = (w_ref, w_het, w_alt, 0) per matched pair, 32-byte rows
    // chunk mode (table == 0): segment j belongs to sample s with seg_off[s] <= j < seg_off[s+1]
    const int32_t *seg_off;     // [S+1]
    const int32_t *mstart;      // [S+1]
    int32_t S;
    int32_t chunk;
    // table mode (table == 1): explicit [begin, end) per segment
    const int32_t *seg_begin;
    const int32_t *seg_end;
    int32_t table;
    int32_t nseg;               // table mode: number of segments; chunk mode: unused (seg_off[S])
    double *part_score;         // [nseg, a_pad]
    int32_t *part_ninfo;        // [nseg, a_pad]
    int32_t a_pad;              // stride * 32
};

__device__ __forceinline__ void score_word(uint64_t word, uint32_t lanebit, int lane, double w_ref, double w_het, double w_alt,
                                           double &s_ref, double &s_het, double &s_alt, int32_t &ninfo, bool skip_hets) {
    const uint32_t lo = uint32_t(word), hi = uint32_t(word >> 32);
    const uint32_t refm = ~(lo | hi), altm = lo & ~hi, hetm = hi & ~lo;
    if (refm & lanebit) s_ref += w_ref;
    if (altm & lanebit) s_alt += w_alt;
    if (!skip_hets) {
        if (hetm != 0u) {                       // warp-uniform: hets are rare (makedb.py:59 code 2)
            if (hetm & lanebit) s_het += w_het;
        }
        ninfo += int32_t((~(lo & hi)) >> lane & 1u);
    } else {
        ninfo += int32_t((refm | altm) >> lane & 1u);   // snpmatch.py:78-79: het -> missing
    }
}

template <bool SKIP_HETS>
__global__ void __launch_bounds__(32 * SC_MAX_WARPS) k_score_segments(const ScoreArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x, warp = threadIdx.y, nwarps = blockDim.y;
    const int seg = blockIdx.x;

    int32_t begin, end;
    if (a.table) {
        if (seg >= a.nseg) return;
        begin = a.seg_begin[seg];
        end = a.seg_end[seg];
    } else {
        if (seg >= a.seg_off[a.S]) return;
        int lo = 0, hi = a.S;                 // last s with seg_off[s] <= seg
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (a.seg_off[mid] <= seg) lo = mid; else hi = mid - 1;
        }
        begin = a.mstart[lo] + (seg - a.seg_off[lo]) * a.chunk;
        end = min(a.mstart[lo + 1], begin + a.chunk);
    }

    const int w_start = blockIdx.y * (SC_WPW * nwarps);
    const int slice_words = min(SC_WPW * nwarps, a.stride - w_start);
    const uint32_t slice_bytes = uint32_t(slice_words) * 8u;
    const int vw = max(0, min(SC_WPW, slice_words - warp * SC_WPW));   // valid words of this warp: 0, 2 or 4

    // shared memory: [STAGES] mbarriers | STAGES x (row tile | weight tile)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    const uint32_t tile_bytes = uint32_t(SC_TILE_ROWS) * uint32_t(SC_WPW * nwarps) * 8u;
    const uint32_t wtile_bytes = SC_TILE_ROWS * 32u;
    unsigned char *stage0 = smem + 128;
    const uint32_t stage_bytes = tile_bytes + wtile_bytes;

    const int n_rows = end - begin;
    const int n_tiles = (n_rows + SC_TILE_ROWS - 1) / SC_TILE_ROWS;

    if (n_tiles > 0) {
        if (warp == 0 && lane == 0) {
            for (int s = 0; s < SC_STAGES; ++s) mbar_init(smem_u32(bars + s), 1u);
            mbar_fence_init();
        }
        __syncthreads();
    }

    auto issue_tile = [&](int t) {          // warp 0, all lanes
        const int st = t % SC_STAGES;
        const int r0 = begin + t * SC_TILE_ROWS;
        const int rows = min(SC_TILE_ROWS, end - r0);
        const uint32_t bar = smem_u32(bars + st);
        unsigned char *dst = stage0 + size_t(st) * stage_bytes;
        if (lane == 0) mbar_arrive_expect_tx(bar, uint32_t(rows) * (slice_bytes + 32u));
        __syncwarp();
        for (int r = lane; r < rows; r += 32) {
            const int64_t row = a.pair_db[r0 + r];
            tma_bulk_g2s(smem_u32(dst + size_t(r) * slice_bytes), a.packed + row * a.stride + w_start, slice_bytes, bar);
        }
        if (lane == 0) tma_bulk_g2s(smem_u32(dst + tile_bytes), a.pair_w + 4 * int64_t(r0), uint32_t(rows) * 32u, bar);
    };

    double s_ref[SC_WPW], s_het[SC_WPW], s_alt[SC_WPW];
    int32_t ninfo[SC_WPW];
#pragma unroll
    for (int j = 0; j < SC_WPW; ++j) { s_ref[j] = 0.0; s_het[j] = 0.0; s_alt[j] = 0.0; ninfo[j] = 0; }
    const uint32_t lanebit = 1u << lane;

    if (warp == 0) {
        for (int t = 0; t < min(n_tiles, SC_STAGES - 1); ++t) issue_tile(t);
    }
    for (int t = 0; t < n_tiles; ++t) {
        if (warp == 0 && t + SC_STAGES - 1 < n_tiles) issue_tile(t + SC_STAGES - 1);
        const int st = t % SC_STAGES;
        mbar_wait(smem_u32(bars + st), uint32_t(t / SC_STAGES) & 1u);
        const int rows = min(SC_TILE_ROWS, n_rows - t * SC_TILE_ROWS);
        const unsigned char *tile = stage0 + size_t(st) * stage_bytes;
        const unsigned char *wt = tile + tile_bytes;
        if (vw > 0) {
            const unsigned char *mine = tile + warp * (SC_WPW * 8);
#pragma unroll 2
            for (int r = 0; r < rows; ++r) {
                const ulonglong2 q0 = *reinterpret_cast<const ulonglong2 *>(mine + size_t(r) * slice_bytes);
                ulonglong2 q1 = make_ulonglong2(~0ull, ~0ull);
                if (vw == SC_WPW) q1 = *reinterpret_cast<const ulonglong2 *>(mine + size_t(r) * slice_bytes + 16);
                const double2 w01 = *reinterpret_cast<const double2 *>(wt + r * 32);
                const double w2 = *reinterpret_cast<const double *>(wt + r * 32 + 16);
                score_word(q0.x, lanebit, lane, w01.x, w01.y, w2, s_ref[0], s_het[0], s_alt[0], ninfo[0], SKIP_HETS);
                score_word(q0.y, lanebit, lane, w01.x, w01.y, w2, s_ref[1], s_het[1], s_alt[1], ninfo[1], SKIP_HETS);
                score_word(q1.x, lanebit, lane, w01.x, w01.y, w2, s_ref[2], s_het[2], s_alt[2], ninfo[2], SKIP_HETS);
                score_word(q1.y, lanebit, lane, w01.x, w01.y, w2, s_ref[3], s_het[3], s_alt[3], ninfo[3], SKIP_HETS);
            }
        }
        __syncthreads();                      // every warp is done with stage st before it is refilled
    }

#pragma unroll
    for (int j = 0; j < SC_WPW; ++j) {
        if (j < vw) {
            const int64_t acc = int64_t(w_start + warp * SC_WPW + j) * 32 + lane;
            const int64_t o = int64_t(seg) * a.a_pad + acc;
            a.part_score[o] = ((0.0 + s_ref[j]) + s_het[j]) + s_alt[j];     // snpmatch.py:84-87
            a.part_ninfo[o] = ninfo[j];
        }
    }
}

static inline void score_launch_shape(int32_t stride, int *nwarps, int *yblocks, size_t *smem_bytes) {
    int nw = (stride + SC_WPW - 1) / SC_WPW;
    if (nw > SC_MAX_WARPS) nw = 8;
    *nwarps = nw;
    *yblocks = (stride + SC_WPW * nw - 1) / (SC_WPW * nw);
    *smem_bytes = 128 + size_t(SC_STAGES) * (size_t(SC_TILE_ROWS) * SC_WPW * nw * 8 + SC_TILE_ROWS * 32);
}

// ---- combine: sequential sum of the segment partials of one sample, in segment order ----------------
// (ScoreList = ScoreList + t_s, snpmatch.py:224-225; TotScoreList, csmatch.py:88-89)
// red row layout [2*n_acc + 2]: score[n_acc] | ninfo[n_acc] as f64 | matched pairs | y>n violations
__global__ void __launch_bounds__(256) k_combine(const double *__restrict__ part_score, const int32_t *__restrict__ part_ninfo,
                                                 int32_t a_pad, int32_t n_acc, const int32_t *__restrict__ seg_off,
                                                 const int32_t *__restrict__ mstart, int32_t nseg_table,
                                                 const int32_t *__restrict__ seg_begin, const int32_t *__restrict__ seg_end,
                                                 double *__restrict__ red) {
    const int s = blockIdx.y;
    const int acc = blockIdx.x * blockDim.x + threadIdx.x;
    int j0, j1;
    if (seg_off) { j0 = seg_off[s]; j1 = seg_off[s + 1]; } else { j0 = 0; j1 = nseg_table; }
    double *row = red + int64_t(s) * (2 * int64_t(n_acc) + 2);
    if (acc < n_acc) {
        double tot = 0.0;
        long long ni = 0;
        for (int j = j0; j < j1; ++j) {
            tot = tot + part_score[int64_t(j) * a_pad + acc];
            ni += part_ninfo[int64_t(j) * a_pad + acc];
        }
        row[acc] = tot;
        row[n_acc + acc] = double(ni);
    }
    if (acc == 0) {
        long long m = 0;
        if (seg_off) m = mstart[s + 1] - mstart[s];
        else for (int j = 0; j < nseg_table; ++j) m += seg_end[j] - seg_begin[j];
        row[2 * n_acc] = double(m);
        row[2 * n_acc + 1] = 0.0;
    }
}

// ---- likelihood epilogue ------------------------------------------------------------------------
// likeliTest (snpmatch.py:40-55)
__device__ __forceinline__ double likeli_test(double n, double y) {
    const double p = 0.99999999;
    if (n == 0.0) return nan("");
    if (y == n) return 1.0;
    if (y > 0.0) {
        const double ps = y / n;
        return y * log(ps / p) + (n - y) * log((1.0 - ps) / (1.0 - p));
    }
    return nan("");
}

__device__ __forceinline__ double block_nanmin(double v, double *s_red /*[32]*/) {
    // nan-ignoring minimum over the block; +inf when every value is nan
    double x = (v == v) ? v : __longlong_as_double(0x7ff0000000000000ll);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, d));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    x = lane < nwarp ? s_red[lane] : __longlong_as_double(0x7ff0000000000000ll);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, d));
    return x;
}

// One CTA per sample row of `red`.  truncate != 0: y = int(score) first (GenotyperOutput.__init__,
// snpmatch.py:96).  amin_mode 0: TopHit = nanmin(L); 1: TopHit = amin.
__global__ void __launch_bounds__(1024) k_epilogue(double *__restrict__ red, int32_t n_acc, int truncate, int amin_mode, double amin,
                                                   int64_t *__restrict__ matches, int64_t *__restrict__ ninfo64,
                                                   double *__restrict__ prob, double *__restrict__ L, double *__restrict__ LR) {
    __shared__ double s_red[32];
    __shared__ int s_viol;
    const int s = blockIdx.x;
    double *row = red + int64_t(s) * (2 * int64_t(n_acc) + 2);
    if (threadIdx.x == 0) s_viol = 0;
    __syncthreads();
    double lmin = nan("");
    int viol = 0;
    for (int acc = threadIdx.x; acc < n_acc; acc += blockDim.x) {
        double y = row[acc];
        const double n = row[n_acc + acc];
        if (truncate) y = double((long long)y);
        const int64_t o = int64_t(s) * n_acc + acc;
        if (matches) matches[o] = (long long)y;
        if (ninfo64) ninfo64[o] = (long long)n;
        prob[o] = n > 0.0 ? y / n : nan("");                  // get_fraction, snpmatch.py:25-28
        if (y > n) ++viol;
        const double l = likeli_test(n, y);
        L[o] = l;
        if (l == l) lmin = (lmin == lmin) ? fmin(lmin, l) : l;
    }
    if (viol) atomicAdd(&s_viol, viol);
    double top = block_nanmin(lmin, s_red);
    if (isinf(top)) top = nan("");                           // all nan (np.nanmin -> nan)
    if (amin_mode) top = amin;
    for (int acc = threadIdx.x; acc < n_acc; acc += blockDim.x) {
        const int64_t o = int64_t(s) * n_acc + acc;
        LR[o] = (top <= 0.0) ? nan("") : L[o] / top;         // get_fraction(L, TopHit)
    }
    __syncthreads();
    if (threadIdx.x == 0) row[2 * n_acc + 1] = double(s_viol);
}

}  // namespace snpm
