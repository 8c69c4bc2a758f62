// A1 — (chrom, pos) join of sample markers against the resident panel.
// Replaces Genotype.get_common_positions (snp_genotype.py:46-68).
//
// Two device algorithms produce the same per-marker result `match_row[i]` (local panel row or -1):
//   * k_join_search   — one binary search per marker inside its chromosome's row range
//                       (n*log2(N) probes; the right tool for low-coverage samples, n << N);
//   * k_join_mergepath — merge-path partition of the two sorted key streams, each CTA merging one
//                       diagonal band from shared memory (reads every panel position once; the
//                       right tool for dense samples and batches).
// An order-preserving compaction (tile counts -> scan -> scatter) then emits the index pairs in
// panel order, the matched weights in matched order, and the inclusive prefix used to cut the
// pairs into samples.
#pragma once
#include "common.cuh"

namespace snpm {

constexpr int JOIN_TILE = 1024;

__device__ __forceinline__ int64_t lower_bound_i32(const int32_t *__restrict__ a, int64_t lo, int64_t hi, int32_t key) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool contains_i64(const int64_t *__restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(a + lo) == key;
}

// status[0] += number of adjacent marker pairs that break the (chromosome, position) order inside a sample
__device__ __forceinline__ void check_order(const int32_t *__restrict__ chrom, const int32_t *__restrict__ pos,
                                            const int64_t *__restrict__ off, int64_t S, int64_t i, int *status) {
    if (i == 0) return;
    const int32_t c1 = chrom[i], c0 = chrom[i - 1];
    if (c1 < 0 || c0 < 0) return;
    const bool bad = (c1 < c0) || (c1 == c0 && pos[i] <= pos[i - 1]);
    if (!bad) return;
    // a new sample may restart the order: is i one of the offsets?
    int64_t lo = 0, hi = S + 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (off[mid] < i) lo = mid + 1; else hi = mid;
    }
    if (lo <= S && off[lo] == i) return;
    atomicAdd(status, 1);
}

// Coarse position index of the panel: bucket b of chromosome c covers positions [b << shift, (b + 1) << shift) and stores the
// first row at or after its start (bucket_off[c] + b; one extra entry per chromosome closes the last bucket).  A marker
// then needs one table read and a short binary search inside ~16 rows that share a few cache lines, instead of
// log2(N) dependent probes spread over the whole position array (which made the search L2-bandwidth bound).
__global__ void __launch_bounds__(256) k_build_buckets(const int32_t *__restrict__ db_pos, const int64_t *__restrict__ chr_regions,
                                                       int32_t n_chr, const int32_t *__restrict__ bucket_off, int shift,
                                                       int32_t *__restrict__ bucket) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= bucket_off[n_chr]) return;
    int c = 0;
    while (c + 1 < n_chr && bucket_off[c + 1] <= i) ++c;
    const int64_t b = i - bucket_off[c];
    const int64_t rs = chr_regions[2 * c], re = chr_regions[2 * c + 1];
    const int64_t key = b << shift;
    bucket[i] = key > 0x7fffffffll ? int32_t(re) : int32_t(lower_bound_i32(db_pos, rs, re, int32_t(key)));
}

// Exact position index of the panel: one bit per base pair of every chromosome (chromosome c starts at bit bm_off[c], a
// multiple of 64) and, per 64-bit word, the row of the first position at or after the word's start.  A marker then needs ONE
// round trip (bitmap word and row word sit at the same index, two independent loads) instead of a bucket read followed by
// four dependent binary-search probes: row = first_row[w] + popc(word & (bit - 1)) when its bit is set.  22 MB for the
// 119 Mbp genome of the 10.7 M-row panel (stays in L2); genomes above 2^31 bp keep the bucket search.
__global__ void __launch_bounds__(256) k_bitmap_set(const int32_t *__restrict__ db_pos, const int64_t *__restrict__ chr_regions, int32_t n_chr,
                                                    const int64_t *__restrict__ bm_off, int64_t n_rows, unsigned long long *__restrict__ bitmap) {
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int c = 0;
    while (c + 1 < n_chr && !(r >= chr_regions[2 * c] && r < chr_regions[2 * c + 1])) ++c;
    if (!(r >= chr_regions[2 * c] && r < chr_regions[2 * c + 1])) return;       // a row outside every chromosome region
    const int32_t p = db_pos[r];
    if (p < 0) return;
    const int64_t bit = bm_off[c] + p;
    atomicOr(bitmap + (bit >> 6), 1ull << (bit & 63));
}
__global__ void __launch_bounds__(256) k_bitmap_rows(const int32_t *__restrict__ db_pos, const int64_t *__restrict__ chr_regions, int32_t n_chr,
                                                     const int64_t *__restrict__ bm_off, int64_t n_words, int32_t *__restrict__ first_row) {
    const int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    int c = 0;
    while (c + 1 < n_chr && bm_off[c + 1] <= w * 64) ++c;
    const int64_t rs = chr_regions[2 * c], re = chr_regions[2 * c + 1];
    const int64_t p0 = w * 64 - bm_off[c];
    first_row[w] = p0 > 0x7fffffffll ? int32_t(re) : int32_t(lower_bound_i32(db_pos, rs, re, int32_t(p0)));
}

__global__ void __launch_bounds__(JOIN_TILE) k_join_search(
        const int32_t *__restrict__ chrom, const int32_t *__restrict__ pos, int64_t n,
        const int64_t *__restrict__ off, int64_t S,
        const int32_t *__restrict__ db_pos, const int64_t *__restrict__ chr_regions, int32_t n_chr,
        const int32_t *__restrict__ bucket, const int32_t *__restrict__ bucket_off, int shift,
        const int64_t *__restrict__ filter, int64_t n_filter, int64_t row0_global,
        int32_t *__restrict__ match_row, int32_t *__restrict__ tile_cnt, int *status, int check,
        const unsigned long long *__restrict__ bitmap, const int32_t *__restrict__ first_row, const int64_t *__restrict__ bm_off,
        const uint32_t *__restrict__ packed_cp) {
    // check == 0: markers are in weight-grouped order (snpm_batch_upload_grouped); the search does not need an order.
    // packed_cp != null: chromosome id and position come in one word (id << 27 | position, id 31 = not in the panel) and are
    // unpacked here instead of by a kernel of their own.
    const int64_t i = int64_t(blockIdx.x) * JOIN_TILE + threadIdx.x;
    int32_t row = -1;
    if (i < n) {
        int32_t c, p;
        if (packed_cp != nullptr) {
            const uint32_t v = packed_cp[i];
            c = (v >> 27) == 31u ? -1 : int32_t(v >> 27);
            p = int32_t(v & 0x7ffffffu);
            if (check && i > 0 && c >= 0) {
                const uint32_t v0 = packed_cp[i - 1];
                if ((v0 >> 27) != 31u && v <= v0) {             // one word compares (chromosome, position) at once
                    int64_t lo = 0, hi = S + 1;                 // a new sample may restart the order: is i one of the offsets?
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (off[mid] < i) lo = mid + 1; else hi = mid;
                    }
                    if (!(lo <= S && off[lo] == i)) atomicAdd(status, 1);
                }
            }
        } else {
            if (check) check_order(chrom, pos, off, S, i, status);
            c = chrom[i];
            p = pos[i];
        }
        if (c >= 0 && c < n_chr) {
            const int32_t b0 = bucket_off[c], nb = bucket_off[c + 1] - b0 - 1;      // buckets of this chromosome
            const int32_t b = p >> shift;
            if (bitmap != nullptr) {
                const int64_t bit = bm_off[c] + p;
                if (p >= 0 && bit < bm_off[c + 1]) {
                    const int64_t w = bit >> 6;
                    const unsigned long long word = __ldg(bitmap + w);
                    const int32_t base = __ldg(first_row + w);
                    const unsigned long long m = 1ull << (bit & 63);
                    if (word & m) {
                        row = base + __popcll(word & (m - 1ull));
                        if (filter && !contains_i64(filter, n_filter, int64_t(row) + row0_global)) row = -1;
                    }
                }
            } else if (p >= 0 && b < nb) {
                const int64_t rs = __ldg(bucket + b0 + b), re = __ldg(bucket + b0 + b + 1);
                const int64_t j = lower_bound_i32(db_pos, rs, re, p);
                if (j < re && __ldg(db_pos + j) == p) {
                    row = int32_t(j);
                    if (filter && !contains_i64(filter, n_filter, j + row0_global)) row = -1;       // an empty list keeps nothing
                }
            }
        }
        match_row[i] = row;
    }
    const int cnt = __syncthreads_count(row >= 0);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = cnt;
}

// ---- merge-path variant -------------------------------------------------------------------------
// Keys: panel side (chromosome c, position) for rows of chromosome c; sample side the same from
// (s_chrom_id, s_pos); both streams ascending in (c, pos) (markers with chromosome -1 are skipped by
// giving them the key of their predecessor's successor... they are handled by the caller: see below).
// One launch per (sample, chromosome) would fragment the work; instead a CTA owns MP_TILE consecutive
// sample markers of ONE chromosome range and the panel slice [lower_bound(first key), upper_bound(last
// key)) that can match them, and merges the two from shared memory.  Panel positions are streamed
// coalesced exactly once over a dense sample; sparse samples make the panel slices long, which is why
// the host picks the search kernel for them.
constexpr int MP_TILE = 1024;          // sample markers per CTA
constexpr int MP_DB_CHUNK = 4096;      // panel positions staged per step

__global__ void __launch_bounds__(256) k_join_mergepath(
        const int32_t *__restrict__ chrom, const int32_t *__restrict__ pos, int64_t n,
        const int64_t *__restrict__ off, int64_t S,
        const int32_t *__restrict__ db_pos, const int64_t *__restrict__ chr_regions, int32_t n_chr,
        const int64_t *__restrict__ filter, int64_t n_filter, int64_t row0_global,
        int32_t *__restrict__ match_row, int32_t *__restrict__ tile_cnt, int *status) {
    __shared__ int32_t s_pos[MP_TILE];
    __shared__ int32_t s_chr[MP_TILE];
    __shared__ int32_t s_db[MP_DB_CHUNK];
    __shared__ int32_t s_cnt;
    __shared__ int32_t s_next;
    __shared__ uint8_t s_brk[MP_TILE];          // marker t starts a new run (chromosome change or order restart)
    const int64_t base = int64_t(blockIdx.x) * MP_TILE;
    const int tile_n = int((n - base) < int64_t(MP_TILE) ? (n - base) : int64_t(MP_TILE));
    if (threadIdx.x == 0) s_cnt = 0;
    for (int t = threadIdx.x; t < tile_n; t += blockDim.x) {
        check_order(chrom, pos, off, S, base + t, status);
        s_pos[t] = pos[base + t];
        s_chr[t] = chrom[base + t];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < tile_n; t += blockDim.x)
        s_brk[t] = t == 0 || s_chr[t] != s_chr[t - 1] || s_pos[t] <= s_pos[t - 1];
    int local_cnt = 0;
    // walk the tile run by run: a run = markers of one chromosome in ascending order (a tile usually holds one run;
    // a chromosome change or the start of the next sample begins another)
    int run_start = 0;
    while (run_start < tile_n) {
        const int32_t c = s_chr[run_start];
        __syncthreads();
        if (threadIdx.x == 0) s_next = tile_n;
        __syncthreads();
        for (int t = threadIdx.x; t < tile_n; t += blockDim.x)
            if (s_brk[t] && t > run_start) atomicMin(&s_next, t);
        __syncthreads();
        const int run_end = s_next;
        if (c < 0 || c >= n_chr) {
            for (int t = run_start + threadIdx.x; t < run_end; t += blockDim.x) match_row[base + t] = -1;
            run_start = run_end;
            continue;
        }
        const int64_t rs = chr_regions[2 * c], re = chr_regions[2 * c + 1];
        const int32_t p_first = s_pos[run_start], p_last = s_pos[run_end - 1];
        // panel slice that can match this run
        const int64_t d0 = lower_bound_i32(db_pos, rs, re, p_first);
        const int64_t d1 = lower_bound_i32(db_pos, d0, re, p_last + 1);
        // thread t owns marker(s) run_start+t...; result accumulates over panel chunks
        for (int t = run_start + threadIdx.x; t < run_end; t += blockDim.x) match_row[base + t] = -1;
        for (int64_t dc = d0; dc < d1; dc += MP_DB_CHUNK) {
            const int chunk_n = int((d1 - dc) < int64_t(MP_DB_CHUNK) ? (d1 - dc) : int64_t(MP_DB_CHUNK));
            __syncthreads();
            for (int t = threadIdx.x; t < chunk_n; t += blockDim.x) s_db[t] = __ldg(db_pos + dc + t);
            __syncthreads();
            const int32_t lo_key = s_db[0], hi_key = s_db[chunk_n - 1];
            // merge-path split: each thread takes an equal share of the (markers-in-range + chunk) diagonal
            // range of markers that can fall into this chunk
            int m0 = run_start, m1 = run_end;
            {   // lower_bound of lo_key / upper_bound of hi_key among the run's markers
                int lo = run_start, hi = run_end;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_pos[mid] < lo_key) lo = mid + 1; else hi = mid; }
                m0 = lo;
                hi = run_end;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_pos[mid] <= hi_key) lo = mid + 1; else hi = mid; }
                m1 = lo;
            }
            const int na = m1 - m0, nb = chunk_n;
            const int total = na + nb;
            const int per = (total + blockDim.x - 1) / blockDim.x;
            const int diag0 = min(total, int(threadIdx.x) * per);
            const int diag1 = min(total, diag0 + per);
            // find (i, j) with i + j = diag0 on the merge path of A = s_pos[m0..m1), B = s_db[0..nb)
            int lo = max(0, diag0 - nb), hi = min(diag0, na);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_pos[m0 + mid] < s_db[diag0 - 1 - mid]) lo = mid + 1; else hi = mid;
            }
            int i = lo, j = diag0 - lo;
            // ties: the panel element goes first (consistent with the strict '<' of the split above), so a
            // marker is matched when the panel element consumed just before it carries the same position
            for (int d = diag0; d < diag1; ++d) {
                const bool take_b = (j < nb) && (i >= na || s_db[j] <= s_pos[m0 + i]);
                if (take_b) {
                    ++j;
                } else {
                    if (j > 0 && s_db[j - 1] == s_pos[m0 + i]) {
                        int32_t row = int32_t(dc + j - 1);
                        if (filter && !contains_i64(filter, n_filter, int64_t(row) + row0_global)) row = -1;
                        match_row[base + m0 + i] = row;
                        if (row >= 0) ++local_cnt;
                    }
                    ++i;
                }
            }
        }
        __syncthreads();
        run_start = run_end;
    }
    atomicAdd(&s_cnt, local_cnt);
    __syncthreads();
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = s_cnt;
}

// ---- compaction ---------------------------------------------------------------------------------
// exclusive scan of `v` over the block (blockDim.x multiple of 32, <= 1024); returns the prefix and
// the block total through *total
__device__ __forceinline__ int block_excl_scan(int v, int *total, int *s_warp /*[33]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarp ? s_warp[lane] : 0;
        int xs = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, xs, d);
            if (lane >= d) xs += y;
        }
        if (lane < nwarp) s_warp[lane] = xs - w;
        if (lane == 31) s_warp[32] = xs;
    }
    __syncthreads();
    const int res = s_warp[warp] + x - v;
    *total = s_warp[32];
    __syncthreads();
    return res;
}

// single CTA: tile_off = exclusive scan of tile_cnt; prefix[n] = total
__global__ void __launch_bounds__(1024) k_scan_tiles(const int32_t *__restrict__ tile_cnt, int64_t n_tiles,
                                                     int32_t *__restrict__ tile_off, int32_t *__restrict__ prefix_last) {
    __shared__ int s_warp[33];
    int carry = 0;
    for (int64_t base = 0; base < n_tiles; base += blockDim.x) {
        const int64_t t = base + threadIdx.x;
        const int v = t < n_tiles ? tile_cnt[t] : 0;
        int total;
        const int ex = block_excl_scan(v, &total, s_warp);
        if (t < n_tiles) tile_off[t] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *prefix_last = carry;
}

// scatter pairs in order; prefix[i] = number of matched markers before i; pair_w = matched weights
__global__ void __launch_bounds__(JOIN_TILE) k_scatter_pairs(
        const int32_t *__restrict__ match_row, int64_t n, const int32_t *__restrict__ tile_off,
        const double *__restrict__ wei, int32_t *__restrict__ prefix, int32_t *__restrict__ pair_db,
        int32_t *__restrict__ pair_s, double *__restrict__ pair_w) {
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
            double4 w4;                        // 32-byte rows: (w_ref, w_alt, w_het, 0) — TMA-copyable tiles
            w4.x = wei[3 * i + 0];
            w4.y = wei[3 * i + 2];
            w4.z = wei[3 * i + 1];
            w4.w = 0.0;
            reinterpret_cast<double4 *>(pair_w)[p] = w4;
        }
    }
}

// dictionary-coded weights -> f64 weights (snpm_batch_upload_indexed)
__global__ void __launch_bounds__(256) k_expand_weights(const uint16_t *__restrict__ idx, const double *__restrict__ table,
                                                        int64_t n3, double *__restrict__ wei) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n3) wei[i] = __ldg(table + idx[i]);
}

// single CTA: matched range of every sample and its number of 1000-row chunks
//   mstart[s] = prefix[off[s]] (mstart[S] = total);  seg_off = exclusive scan of ceil(m_s / chunk)
__global__ void __launch_bounds__(1024) k_sample_ranges(const int32_t *__restrict__ prefix, const int64_t *__restrict__ off,
                                                        int64_t S, int32_t chunk, int32_t *__restrict__ mstart,
                                                        int32_t *__restrict__ seg_off) {
    __shared__ int s_warp[33];
    for (int64_t s = threadIdx.x; s <= S; s += blockDim.x) mstart[s] = prefix[off[s]];
    __syncthreads();
    int carry = 0;
    for (int64_t base = 0; base < S; base += blockDim.x) {
        const int64_t s = base + threadIdx.x;
        int v = 0;
        if (s < S) v = (mstart[s + 1] - mstart[s] + chunk - 1) / chunk;
        int total;
        const int ex = block_excl_scan(v, &total, s_warp);
        if (s < S) seg_off[s] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) seg_off[S] = carry;
}

}  // namespace snpm
