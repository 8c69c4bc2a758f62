// Device-side grouping of a batch's matched markers by weight triple — what snpm_group_markers did on the host in round 1.
//
// The host uploads what a parser has in hand (parsers.py:141-157): markers in position order as (chromosome id, position)
// words plus, per marker, three dictionary codes (ref, het, alt) into a table of distinct weight values (for a VCF the code
// IS the integer PL and the table is exp(-PL/10)).  After the (chrom, pos) join the kernels here
//   1. give every matched pair a sort key  [called class | code of the called class | code of the slow class | code of the
//      fast class]  (k_scatter_pairs_coded) — the hierarchical order that makes every class weight change as rarely as
//      possible along a sample (the counting kernel reads a class counter out only when THAT class's weight changes),
//   2. sort the pairs of every sample by that key with a stable, segmented LSD radix sort (k_radix_hist / _scan / _scatter:
//      digits of up to 11 bits, tile histograms, warp-aggregated ranks via match.any) — deterministic, position order is kept
//      inside a group,
//   3. mark, per block of 16 sorted rows, where each class weight changes (k_group_masks).
// Replaces the per-sample work of Genotyper.genotyper's chunk loop set-up (snpmatch.py:218-227) in the grouped formulation;
// results do not depend on the order (counts are order-free), only the speed does.
#pragma once
#include "common.cuh"
#include "join.cuh"

namespace snpm {

constexpr int RS_THREADS = 256;
constexpr int RS_PER_THREAD = 8;
constexpr int RS_TILE = RS_THREADS * RS_PER_THREAD;       // pairs per sort tile
constexpr int RS_MAX_BITS = 11;                            // digit width (2048 bins: 8 warps x 2048 x u16 = 32 KB of shared memory)

// class indices of the scoring kernels: 0 ref, 1 alt, 2 het.  Called class c -> (slow, fast) = the remaining classes, the one
// whose weight takes fewer distinct values first (het for homozygous calls: 3*DP-like PLs; ref for het calls).
__host__ __device__ __forceinline__ int gs_slow_class(int c) { return c == 2 ? 0 : 2; }
__host__ __device__ __forceinline__ int gs_fast_class(int c) { return c == 1 ? 0 : 1; }
// bit offset of class `which`'s code inside a key whose called class is c (b = bits per code)
__host__ __device__ __forceinline__ int gs_field_shift(int c, int which, int b) {
    return which == c ? 2 * b : (which == gs_slow_class(c) ? b : 0);
}
template <typename KeyT>
__host__ __device__ __forceinline__ uint32_t gs_code(KeyT key, int which, int b) {
    const int c = int(key >> (3 * b)) & 3;
    return uint32_t(key >> gs_field_shift(c, which, b)) & ((1u << b) - 1u);
}

// as k_scatter_pairs (join.cuh), the payload of a pair being its sort key and its own index (the sort's initial permutation)
template <typename KeyT>
__global__ void __launch_bounds__(JOIN_TILE) k_scatter_pairs_coded(
        const int32_t *__restrict__ match_row, int64_t n, const int32_t *__restrict__ tile_off, const uint16_t *__restrict__ codes,
        const double *__restrict__ wtable, int32_t n_table, int32_t code_bits, int32_t *__restrict__ prefix,
        int32_t *__restrict__ pair_db, int32_t *__restrict__ pair_s, KeyT *__restrict__ key, uint32_t *__restrict__ idx, int *status) {
    __shared__ int s_warp[33];
    const int64_t i = int64_t(blockIdx.x) * JOIN_TILE + threadIdx.x;
    const int32_t row = i < n ? match_row[i] : -1;
    const int flag = row >= 0;
    int total;
    const int ex = block_excl_scan(flag, &total, s_warp);
    if (i < n) {
        const int32_t p = tile_off[blockIdx.x] + ex;
        prefix[i] = p;
        if (flag) {
            pair_db[p] = row;
            pair_s[p] = int32_t(i);
            uint32_t cd[3] = {codes[3 * i], codes[3 * i + 2], codes[3 * i + 1]};      // wei columns are (ref, het, alt); classes (ref, alt, het)
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (int32_t(cd[k]) >= n_table) { bad = true; cd[k] = 0; }
            if (bad) atomicAdd(status + 4, 1);             // a code outside the table: reported at wait / fetch
            const double w0 = __ldg(wtable + cd[0]), w1 = __ldg(wtable + cd[1]), w2 = __ldg(wtable + cd[2]);
            int c;                                          // called class: the weight that is 1.0, else the largest
            if (w0 == 1.0) c = 0; else if (w1 == 1.0) c = 1; else if (w2 == 1.0) c = 2;
            else { c = 0; if (w1 > w0) c = 1; if (w2 > (c ? w1 : w0)) c = 2; }
            const int b = code_bits;
            key[p] = (KeyT(c) << (3 * b)) | (KeyT(cd[c]) << (2 * b)) | (KeyT(cd[gs_slow_class(c)]) << b) | KeyT(cd[gs_fast_class(c)]);
            idx[p] = uint32_t(p);
        }
    }
}

// tile t of the sort covers pairs [mstart[s] + lt * RS_TILE, ...) of sample s = tile_sample[t], lt = t - tile_first[s]; the
// tile layout comes from the host's upper bound (markers per sample), so a tile may be empty
__device__ __forceinline__ void rs_tile_range(const int32_t *__restrict__ mstart, const int32_t *__restrict__ tile_sample,
                                              const int32_t *__restrict__ tile_first, int t, int *begin, int *end, int *sample) {
    const int s = tile_sample[t];
    const int lt = t - tile_first[s];
    const int b0 = mstart[s], b1 = mstart[s + 1];
    *sample = s;
    *begin = min(b1, b0 + lt * RS_TILE);
    *end = min(b1, *begin + RS_TILE);
}

// digit histogram of one tile -> tile_hist[t][0..bins)
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const KeyT *__restrict__ key, const int32_t *__restrict__ mstart,
                                                          const int32_t *__restrict__ tile_sample, const int32_t *__restrict__ tile_first,
                                                          int shift, int bits, uint32_t *__restrict__ tile_hist) {
    extern __shared__ uint32_t rs_h[];
    const int bins = 1 << bits;
    for (int d = threadIdx.x; d < bins; d += RS_THREADS) rs_h[d] = 0u;
    __syncthreads();
    int begin, end, s;
    rs_tile_range(mstart, tile_sample, tile_first, blockIdx.x, &begin, &end, &s);
    const uint32_t dmask = uint32_t(bins - 1);
    const int lane = threadIdx.x & 31;
    for (int i0 = begin + (threadIdx.x & ~31); i0 < end; i0 += RS_THREADS) {
        const int i = i0 + lane;
        const uint32_t d = i < end ? uint32_t(key[i] >> shift) & dmask : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (i < end && lane == __ffs(peers) - 1) atomicAdd(rs_h + d, uint32_t(__popc(peers)));
    }
    __syncthreads();
    uint32_t *out = tile_hist + size_t(blockIdx.x) * bins;
    for (int d = threadIdx.x; d < bins; d += RS_THREADS) out[d] = rs_h[d];
}

// one CTA per sample: tile_hist[t][d] -> exclusive offset of (digit d, tile t) inside the sample's sorted range
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t *__restrict__ tile_hist, const int32_t *__restrict__ tile_first, int bits) {
    __shared__ int s_warp[33];
    const int bins = 1 << bits;
    const int s = blockIdx.x;
    const int t0 = tile_first[s], t1 = tile_first[s + 1];
    // thread -> `per` consecutive digits
    const int per = (bins + int(blockDim.x) - 1) / int(blockDim.x);
    const int d0 = threadIdx.x * per;
    uint32_t run[2] = {0u, 0u};                      // per <= 2 (2048 bins, 1024 threads)
    for (int t = t0; t < t1; ++t) {
        uint32_t *h = tile_hist + size_t(t) * bins;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k < per && d0 + k < bins) {
                const uint32_t v = h[d0 + k];
                h[d0 + k] = run[k];
                run[k] += v;
            }
        }
    }
    int total;
    const int mine = int(run[0] + (per > 1 ? run[1] : 0u));
    const int ex = block_excl_scan(mine, &total, s_warp);
    const uint32_t base0 = uint32_t(ex), base1 = uint32_t(ex) + run[0];
    for (int t = t0; t < t1; ++t) {
        uint32_t *h = tile_hist + size_t(t) * bins;
        if (d0 < bins) h[d0] += base0;
        if (per > 1 && d0 + 1 < bins) h[d0 + 1] += base1;
    }
}

// stable scatter of one tile.  Warp w owns pairs [256 w, 256 w + 256) of the tile and takes them 32 at a time in order, so
// ranks inside a digit follow the pair order: rank = (pairs of that digit in earlier warps) + (earlier rounds of this warp) +
// (lower lanes of this round, match.any).  LAST: the payload index is resolved to the pair itself (panel row, marker index).
template <typename KeyT, bool LAST>
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const KeyT *__restrict__ key_in, const uint32_t *__restrict__ idx_in,
                                                             KeyT *__restrict__ key_out, uint32_t *__restrict__ idx_out,
                                                             const int32_t *__restrict__ pair_db_in, const int32_t *__restrict__ pair_s_in,
                                                             int32_t *__restrict__ pair_db_out, int32_t *__restrict__ pair_s_out,
                                                             const int32_t *__restrict__ mstart, const int32_t *__restrict__ tile_sample,
                                                             const int32_t *__restrict__ tile_first, const uint32_t *__restrict__ tile_off,
                                                             int shift, int bits) {
    extern __shared__ uint32_t rs_sm[];
    const int bins = 1 << bits;
    uint16_t *whist = reinterpret_cast<uint16_t *>(rs_sm);          // [8][bins]
    constexpr int NW = RS_THREADS / 32;
    for (int k = threadIdx.x; k < NW * bins / 2; k += RS_THREADS) rs_sm[k] = 0u;
    __syncthreads();
    int begin, end, s;
    rs_tile_range(mstart, tile_sample, tile_first, blockIdx.x, &begin, &end, &s);
    if (begin >= end) return;
    const uint32_t dmask = uint32_t(bins - 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint16_t *mine = whist + size_t(warp) * bins;
    KeyT k[RS_PER_THREAD];
    uint32_t v[RS_PER_THREAD], rank[RS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        const bool on = i < end;
        k[e] = on ? key_in[i] : KeyT(0);
        v[e] = on ? idx_in[i] : 0u;
        const uint32_t d = on ? uint32_t(k[e] >> shift) & dmask : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t prior = 0u;
        if (on) prior = mine[d];
        __syncwarp();
        if (on && lane == __ffs(peers) - 1) mine[d] = uint16_t(prior + __popc(peers));
        __syncwarp();
        rank[e] = prior + uint32_t(__popc(peers & lt_mask));
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps (in place)
    for (int d = threadIdx.x; d < bins; d += RS_THREADS) {
        uint32_t run = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t c = whist[size_t(w) * bins + d];
            whist[size_t(w) * bins + d] = uint16_t(run);
            run += c;
        }
    }
    __syncthreads();
    const uint32_t *off = tile_off + size_t(blockIdx.x) * bins;
    const int base = mstart[s];
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        if (i < end) {
            const uint32_t d = uint32_t(k[e] >> shift) & dmask;
            const int o = base + int(off[d]) + int(mine[d]) + int(rank[e]);
            key_out[o] = k[e];
            if (LAST) {
                pair_db_out[o] = pair_db_in[v[e]];
                pair_s_out[o] = pair_s_in[v[e]];
            } else {
                idx_out[o] = v[e];
            }
        }
    }
}

// Per block of 16 sorted rows of a sample (blocks counted from the sample's first pair; a segment of `chunk` rows is chunk/16
// blocks): three 16-bit masks (ref | alt << 16 | het << 32), bit k set <=> the weight code of that class at row 16 b + k
// differs from the row before it.  The first row of a segment is never marked (the kernel loads its weights afresh).
// grid (ceil(max pairs of a sample / 256), S).
template <typename KeyT>
__global__ void __launch_bounds__(256) k_group_masks(const KeyT *__restrict__ key, const int32_t *__restrict__ mstart,
                                                     const int32_t *__restrict__ seg_off, int32_t chunk, int32_t code_bits,
                                                     unsigned long long *__restrict__ blk_chg) {
    const int s = blockIdx.y;
    const int b0 = mstart[s], m = mstart[s + 1] - b0;
    const int r = blockIdx.x * 256 + threadIdx.x;
    if ((r & ~31) >= m) return;                       // whole warp past the end
    uint32_t f[3] = {0u, 0u, 0u};
    if (r < m && r % chunk != 0) {
        const KeyT k1 = key[b0 + r], k0 = key[b0 + r - 1];
        const int b = code_bits;
        const int c1 = int(k1 >> (3 * b)) & 3, c0 = int(k0 >> (3 * b)) & 3;
        if (c1 != c0) {
            f[0] = f[1] = f[2] = 1u;
        } else {
            const KeyT d = k1 ^ k0;
            const uint32_t fm = (1u << b) - 1u;
#pragma unroll
            for (int w = 0; w < 3; ++w) f[w] = (uint32_t(d >> gs_field_shift(c1, w, b)) & fm) != 0u;
        }
    }
    const uint32_t b_ref = __ballot_sync(0xffffffffu, f[0]), b_alt = __ballot_sync(0xffffffffu, f[1]), b_het = __ballot_sync(0xffffffffu, f[2]);
    const int lane = threadIdx.x & 31;
    if ((lane & 15) == 0 && r < m) {
        const int sh = lane;                          // 0 or 16
        const unsigned long long v = (unsigned long long)((b_ref >> sh) & 0xffffu) | ((unsigned long long)((b_alt >> sh) & 0xffffu) << 16) |
                                     ((unsigned long long)((b_het >> sh) & 0xffffu) << 32);
        blk_chg[size_t(seg_off[s]) * size_t(chunk / 16) + size_t(r >> 4)] = v;
    }
}

}  // namespace snpm
