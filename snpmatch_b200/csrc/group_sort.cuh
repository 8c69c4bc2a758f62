// Device-side grouping of a batch's matched markers by weight triple — what snpm_group_markers did on the host in round 1.
//
// The host uploads what a parser has in hand (parsers.py:141-157): markers in position order as (chromosome id, position)
// words plus, per marker, three dictionary codes (ref, het, alt) into a table of distinct weight values (for a VCF the code
// IS the integer PL and the table is exp(-PL/10)).  After the (chrom, pos) join the kernels here
//   1. give every matched pair a sort key  [called class | code of the called class | code of the slow class | code of the
//      fast class]  (k_scatter_pairs_coded) — the hierarchical order that makes every class weight change as rarely as
//      possible along a sample (the counting kernel reads a class counter out only when THAT class's weight changes),
//   2. sort the pairs of every sample by that key with a stable, segmented LSD radix sort (k_radix_hist once, then one
//      k_radix_pass per digit of up to 11 bits: offsets from the tile histograms, warp-aggregated ranks via match.any, scatter,
//      next digit's histogram on the way) — deterministic, position order is kept inside a group,
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

// lanes of the warp that hold the same `bits`-bit digit as the caller (valid lanes only): one ballot per digit bit.  (match.any
// does the same in one instruction but its cost grows with the number of distinct values in the warp — up to 32 here.)
__device__ __forceinline__ uint32_t rs_peers(uint32_t d, bool valid, int bits) {
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
    for (int b = 0; b < bits; ++b) {
        const uint32_t bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// Tile t of the sort covers pairs [begin, end) of sample tile_sample[t]: the lt-th RS_TILE pairs of the sample's matched range
// (lt = t - tile_first[sample]).  The tile layout comes from the host's upper bound (markers per sample), so a tile may be empty.
// One thread per tile; runs once per step behind k_sample_ranges so that the sort kernels find their range with ONE load
// instead of a chain of three dependent ones.
__global__ void __launch_bounds__(256) k_tile_ranges(const int32_t *__restrict__ mstart, const int32_t *__restrict__ tile_sample,
                                                     const int32_t *__restrict__ tile_first, int32_t n_tiles, int2 *__restrict__ range) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int s = tile_sample[t];
    const int lt = t - tile_first[s];
    const int b0 = mstart[s], b1 = mstart[s + 1];
    const int begin = min(b1, b0 + lt * RS_TILE);
    range[t] = make_int2(begin, min(b1, begin + RS_TILE));
}

// digit histogram of one tile -> tile_hist[t][0..bins)  (first pass only: later passes get theirs from the scatter before them)
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const KeyT *__restrict__ key, const int2 *__restrict__ range,
                                                          int shift, int bits, uint32_t *__restrict__ tile_hist) {
    extern __shared__ uint32_t rs_h[];
    const int bins = 1 << bits;
    const int2 rg = range[blockIdx.x];
    for (int d = threadIdx.x; d < bins; d += RS_THREADS) rs_h[d] = 0u;
    __syncthreads();
    const int begin = rg.x, end = rg.y;
    const uint32_t dmask = uint32_t(bins - 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KeyT k[RS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {                      // all loads first: the tile is one DRAM round trip, not eight
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        k[e] = i < end ? key[i] : KeyT(0);
    }
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        const uint32_t d = uint32_t(k[e] >> shift) & dmask;
        const uint32_t peers = rs_peers(d, i < end, bits);
        if (i < end && lane == __ffs(peers) - 1) atomicAdd(rs_h + d, uint32_t(__popc(peers)));
    }
    __syncthreads();
    uint32_t *out = tile_hist + size_t(blockIdx.x) * bins;
    for (int d = threadIdx.x; d < bins; d += RS_THREADS) out[d] = rs_h[d];
}

// One pass of the stable segmented LSD radix sort, one CTA per tile, three steps in one kernel:
//   offsets  where the tile's pairs of digit d go inside the sample's range = (pairs of smaller digits in the whole sample) +
//            (pairs of digit d in earlier tiles of the sample), from the per-tile digit counts of ALL tiles of the sample
//            (a few tens of KB out of L2 per CTA: cheaper than a separate scan kernel between two dependent launches);
//   ranks    warp w owns pairs [256 w, 256 w + 256) of the tile and takes them 32 at a time in order, so ranks inside a digit
//            follow the pair order: (pairs of that digit in earlier warps) + (earlier rounds of this warp) + (lower lanes of
//            this round, match.any);
//   scatter  key + payload to their place; the NEXT pass's per-tile digit counts are accumulated on the way (atomics into a
//            zeroed table, indexed by the tile the pair lands in).  LAST: the payload index is resolved to the pair itself
//            (panel row, marker index).
// Dynamic shared memory: bins * 20 bytes (8 warp histograms of u16 + one u32 offset per digit).
template <typename KeyT, bool LAST>
__global__ void __launch_bounds__(RS_THREADS, 4) k_radix_pass(const KeyT *__restrict__ key_in, const uint32_t *__restrict__ idx_in,
                                                             KeyT *__restrict__ key_out, uint32_t *__restrict__ idx_out,
                                                             const int32_t *__restrict__ pair_db_in, const int32_t *__restrict__ pair_s_in,
                                                             int32_t *__restrict__ pair_db_out, int32_t *__restrict__ pair_s_out,
                                                             const int32_t *__restrict__ mstart, const int32_t *__restrict__ tile_sample,
                                                             const int32_t *__restrict__ tile_first, const int2 *__restrict__ range,
                                                             const uint32_t *__restrict__ hist_in, uint32_t *__restrict__ hist_out,
                                                             int shift, int bits, int next_shift, int next_bits) {
    extern __shared__ uint32_t rs_sm[];
    __shared__ int s_warp[33];
    const int bins = 1 << bits;
    constexpr int NW = RS_THREADS / 32;
    uint16_t *whist = reinterpret_cast<uint16_t *>(rs_sm);          // [NW][bins]
    uint32_t *off = rs_sm + NW * bins / 2;                          // [bins]
    const int t = blockIdx.x;
    const int2 rg = range[t];
    const int begin = rg.x, end = rg.y;
    if (begin >= end) return;                                       // the whole CTA: an empty tile
    const int s = tile_sample[t];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the tile's pairs: loads issued before anything depends on them
    KeyT k[RS_PER_THREAD];
    uint32_t v[RS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        const bool on = i < end;
        k[e] = on ? key_in[i] : KeyT(0);
        v[e] = on ? idx_in[i] : 0u;
    }
    for (int j = threadIdx.x; j < NW * bins / 2; j += RS_THREADS) rs_sm[j] = 0u;
    const int t0 = tile_first[s], t1 = tile_first[s + 1], base = mstart[s];
    // offsets: thread -> `per` consecutive digits
    {
        const int per = (bins + RS_THREADS - 1) / RS_THREADS;       // 1, 2, 4 or 8
        const int d0 = threadIdx.x * per;
        uint32_t total[8], before[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) total[j] = before[j] = 0u;
        if (d0 < bins) {
            constexpr int U = 8;                                    // tiles per batch of independent loads (the loop is latency bound)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < per) {
                    for (int tt = t0; tt < t1; tt += 2 * U) {
                        uint32_t c[2 * U];
#pragma unroll
                        for (int u = 0; u < 2 * U; ++u) c[u] = tt + u < t1 ? __ldg(hist_in + size_t(tt + u) * bins + d0 + j) : 0u;
#pragma unroll
                        for (int u = 0; u < 2 * U; ++u) {
                            total[j] += c[u];
                            if (tt + u < t) before[j] += c[u];
                        }
                    }
                }
            }
        }
        uint32_t mine = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) mine += total[j];
        int all;
        uint32_t run = uint32_t(block_excl_scan(int(mine), &all, s_warp));
        if (d0 < bins) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < per) {
                    off[d0 + j] = run + before[j];
                    run += total[j];
                }
            }
        }
    }
    __syncthreads();
    const uint32_t dmask = uint32_t(bins - 1);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint16_t *mine_h = whist + size_t(warp) * bins;
    uint32_t rank[RS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        const bool on = i < end;
        const uint32_t d = uint32_t(k[e] >> shift) & dmask;
        const uint32_t peers = rs_peers(d, on, bits);
        uint32_t prior = 0u;
        if (on) prior = mine_h[d];
        __syncwarp();
        if (on && lane == __ffs(peers) - 1) mine_h[d] = uint16_t(prior + __popc(peers));
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
    const uint32_t nmask = next_bits > 0 ? uint32_t((1 << next_bits) - 1) : 0u;
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        if (i < end) {
            const uint32_t d = uint32_t(k[e] >> shift) & dmask;
            const int local = int(off[d]) + int(mine_h[d]) + int(rank[e]);
            const int o = base + local;
            key_out[o] = k[e];
            if (LAST) {
                pair_db_out[o] = pair_db_in[v[e]];
                pair_s_out[o] = pair_s_in[v[e]];
            } else {
                idx_out[o] = v[e];
            }
        }
        if (!LAST && hist_out != nullptr && i < end) {
            // next pass's digit counts, booked on the tile the pair lands in (fire-and-forget atomics into a zeroed table)
            const uint32_t d = uint32_t(k[e] >> shift) & dmask;
            const int local = int(off[d]) + int(mine_h[d]) + int(rank[e]);
            atomicAdd(hist_out + ((size_t(t0) + size_t(local / RS_TILE)) << next_bits) + (uint32_t(k[e] >> next_shift) & nmask), 1u);
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
