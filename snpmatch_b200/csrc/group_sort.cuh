// Device-side grouping of a batch's matched markers by weight triple — what snpm_group_markers did on the host in round 1.
//
// The host uploads what a parser has in hand (parsers.py:141-157): markers in position order as (chromosome id, position)
// words plus, per marker, three dictionary codes (ref, het, alt) into a table of distinct weight values (for a VCF the code
// IS the integer PL and the table is exp(-PL/10)).  After the (chrom, pos) join the kernels here
//   1. give every matched pair a key  [called class | code of the called class | code of the slow class | code of the fast
//      class]  — ordering by it is the hierarchical order that makes every class weight change as rarely as possible along a
//      sample (the counting kernel reads a class counter out only when THAT class's weight changes) — and collect the
//      distinct keys of every sample in a small hash table (k_scatter_pairs_coded),
//   2. sort each sample's keys (a few hundred): the rank of a key is the dense id of its group; the group's three weights go
//      into a table (k_group_rank),
//   3. look the id of every pair up and count ids per tile of 2048 pairs (k_group_ids),
//   4. move the panel rows of the pairs, ONCE, from position order to (id, position) order with a stable partition
//      (k_group_place: offsets from the tile counts, ranks in pair order, scatter) — deterministic,
//   5. mark, per block of 16 grouped rows, where each class weight changes and which group the block starts in
//      (k_group_marks), from the group table alone.
// Steps 3-5 run either as kernels over tiles of 2048 pairs (k_group_ids, k_group_scan, k_group_place, k_group_marks: any sample
// size) or, for samples of up to 65 535 pairs, as ONE kernel with a CTA per sample that keeps the sample's counters in shared
// memory (k_group_sample); api.cu picks by batch shape.  Both produce the same order.
// A radix sort over the ~30-bit key did the same in three scatter passes of key + permutation (0.20 ms for 2.9 M pairs against
// a 0.27 ms scoring kernel: scattered 4-byte stores are the cost, one LSU transaction each); the dense ids need one pass over
// one array.  Replaces the per-sample work of Genotyper.genotyper's chunk loop set-up (snpmatch.py:218-227) in the grouped
// formulation; results do not depend on the order (counts are order-free), only the speed does.
#pragma once
#include "common.cuh"
#include "join.cuh"

namespace snpm {

constexpr int RS_THREADS = 256;
constexpr int RS_PER_THREAD = 8;
constexpr int RS_TILE = RS_THREADS * RS_PER_THREAD;       // pairs per tile of the partition pass

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

// ---- dense group ids: one open-addressing hash table of sort keys per sample -------------------------------------------
// A sample carries a few hundred distinct weight triples (740 among the 50 000 markers of a synthetic PL sample, 796 among
// the 7545 of 701_501.filter.vcf), so its pairs can be ordered with ONE stable partition pass over dense ids instead of a
// radix sort over the ~30-bit key: the keys of a sample are collected in its table while the pairs are compacted
// (gh_insert), sorted (k_group_rank: the id of a key is its rank, so ordering by id is ordering by key), looked up per pair
// (k_group_ids) and the pairs placed (k_group_place).  A sample with more than GH_MAX_GROUPS distinct triples is flagged
// (its groups would be a handful of rows each: the order-exact kernel is the right tool) and re-scored by the caller.
constexpr int GH_SLOTS = 4096;            // slots per sample
constexpr int GH_MAX_GROUPS = 2048;       // dense ids per sample: one 11-bit digit

__device__ __forceinline__ uint32_t gh_hash(unsigned long long key) { return uint32_t((key * 0x9E3779B97F4A7C15ull) >> 52); }
// a slot holds key + 1 (0 = empty); slots only ever change from empty to taken
__device__ __forceinline__ void gh_insert(unsigned long long *__restrict__ table, unsigned long long key, int *overflow) {
    const unsigned long long want = key + 1ull;
    uint32_t h = gh_hash(key);
    for (int probe = 0; probe < GH_SLOTS; ++probe) {
        unsigned long long cur = __ldcg(table + h);
        if (cur == 0ull) cur = atomicCAS(table + h, 0ull, want);
        if (cur == 0ull || cur == want) return;
        h = (h + 1u) & uint32_t(GH_SLOTS - 1);
    }
    atomicExch(overflow, 1);
}
__device__ __forceinline__ uint32_t gh_find(const unsigned long long *__restrict__ table, unsigned long long key) {
    const unsigned long long want = key + 1ull;
    uint32_t h = gh_hash(key);
    for (int probe = 0; probe < GH_SLOTS; ++probe) {
        const unsigned long long cur = __ldg(table + h);
        if (cur == want || cur == 0ull) break;
        h = (h + 1u) & uint32_t(GH_SLOTS - 1);
    }
    return h;
}

// as k_scatter_pairs (join.cuh), the payload of a pair being its key, which also goes into its sample's key table.  A CTA of
// SC_THREADS threads takes one tile of JOIN_TILE markers, four per thread (marker j * SC_THREADS + thread of the tile: all
// loads of a thread are issued before the first is used), ranks the matched ones with warp ballots and ONE block barrier, and
// inserts every distinct (sample, key) of the tile once: lanes of a warp that hold the same combination elect one
// (match.any), which asks a shared-memory table of the CTA before it probes the sample's table in global memory (an L2 round
// trip behind an atomicCAS: 40 % of the stall samples of the first version, a 1024-thread CTA with one marker per thread and
// three barriers: 82 us per 3.2 M markers).
constexpr int SC_THREADS = 256;
constexpr int SC_PER = JOIN_TILE / SC_THREADS;
constexpr int SC_SLOTS = 1024;
template <typename KeyT>
__global__ void __launch_bounds__(SC_THREADS) k_scatter_pairs_coded(
        const int32_t *__restrict__ match_row, int64_t n, const int32_t *__restrict__ tile_off, const uint16_t *__restrict__ codes,
        const double *__restrict__ wtable, int32_t n_table, int32_t code_bits, int32_t *__restrict__ prefix,
        int32_t *__restrict__ pair_db, int32_t *__restrict__ pair_s, KeyT *__restrict__ key, int *status,
        const int64_t *__restrict__ off, int64_t S, unsigned long long *__restrict__ hash, int *__restrict__ overflow,
        const uint32_t *__restrict__ codes32) {
    constexpr int NW = SC_THREADS / 32;
    __shared__ int s_cnt[SC_PER][NW];
    __shared__ int64_t s_first_sm, s_first_end;
    __shared__ unsigned long long s_seen[SC_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = int64_t(blockIdx.x) * JOIN_TILE;
    int32_t row[SC_PER];
    uint32_t cw[SC_PER];
#pragma unroll
    for (int j = 0; j < SC_PER; ++j) {
        const int64_t i = i0 + j * SC_THREADS + threadIdx.x;
        row[j] = i < n ? match_row[i] : -1;
    }
#pragma unroll
    for (int j = 0; j < SC_PER; ++j) {
        const int64_t i = i0 + j * SC_THREADS + threadIdx.x;
        cw[j] = (codes32 != nullptr && i < n) ? codes32[i] : 0u;
    }
    const int32_t base0 = tile_off[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SC_SLOTS / SC_THREADS; ++j) s_seen[j * SC_THREADS + threadIdx.x] = 0ull;
    uint32_t bal[SC_PER];
#pragma unroll
    for (int j = 0; j < SC_PER; ++j) {
        bal[j] = __ballot_sync(0xffffffffu, row[j] >= 0);
        if (lane == 0) s_cnt[j][warp] = __popc(bal[j]);
    }
    if (threadIdx.x == 0) {                         // sample of the tile's first marker, and where that sample ends
        int64_t lo = 0, hi = S;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (off[mid] <= i0) lo = mid; else hi = mid;
        }
        s_first_sm = lo;
        s_first_end = off[lo + 1];
    }
    __syncthreads();
    const int64_t s_first = s_first_sm, first_end = s_first_end;
    int32_t base = base0;
#pragma unroll
    for (int j = 0; j < SC_PER; ++j) {
        const int64_t i = i0 + j * SC_THREADS + threadIdx.x;
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int c = s_cnt[j][w];
            before += w < warp ? c : 0;
            total += c;
        }
        const int32_t p = base + before + __popc(bal[j] & ((1u << lane) - 1u));
        base += total;
        unsigned long long ins = ~0ull, fold = 0ull;    // the key to insert, or nothing
        bool solo = false;
        int64_t smp = 0;
        if (i < n) {
            prefix[i] = p;
            if (row[j] >= 0) {
                pair_db[p] = row[j];
                pair_s[p] = int32_t(i);
                uint32_t cd[3];                             // wei columns are (ref, het, alt); classes (ref, alt, het)
                if (codes32 != nullptr) {                   // three 10-bit codes in one word: ref | het << 10 | alt << 20
                    const uint32_t v = cw[j];
                    cd[0] = v & 1023u; cd[2] = (v >> 10) & 1023u; cd[1] = (v >> 20) & 1023u;
                } else {
                    cd[0] = codes[3 * i]; cd[2] = codes[3 * i + 1]; cd[1] = codes[3 * i + 2];
                }
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
                const KeyT kk = (KeyT(c) << (3 * b)) | (KeyT(cd[c]) << (2 * b)) | (KeyT(cd[gs_slow_class(c)]) << b) | KeyT(cd[gs_fast_class(c)]);
                key[p] = kk;
                int64_t lo = s_first;                           // sample of marker i: the last offset <= i (a tile lies inside one sample
                if (i >= first_end) {                           // almost always: no load then; else it spans few samples)
                    ++lo;
                    while (lo + 1 < S && off[lo + 1] <= i) ++lo;
                }
                smp = lo;
                // the sample is folded into the word that is matched (keys have at most 50 bits) when it is close enough to the tile's first
                ins = (unsigned long long)(kk);
                if (lo - s_first < 8192) fold = (unsigned long long)(lo - s_first) << 50;
                else solo = true;
            }
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, solo ? (0xffffull << 48 | (unsigned long long)(threadIdx.x)) : (ins | fold));
        if (ins != ~0ull && (solo || lane == __ffs(peers) - 1)) {
            bool fresh = true;                              // not inserted by this CTA yet (or its table cannot tell)
            if (!solo) {
                const unsigned long long want = (ins | fold) + 1ull;
                uint32_t h = uint32_t((want * 0x9E3779B97F4A7C15ull) >> 54);
                for (int probe = 0; probe < 16; ++probe) {
                    const unsigned long long cur = atomicCAS(s_seen + h, 0ull, want);
                    if (cur == 0ull) break;
                    if (cur == want) { fresh = false; break; }
                    h = (h + 1u) & uint32_t(SC_SLOTS - 1);
                }
            }
            if (fresh) gh_insert(hash + size_t(smp) * GH_SLOTS, ins, overflow + smp);
        }
    }
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


// One CTA per sample: the sample's distinct keys (its hash table, compacted) are sorted in shared memory (bitonic over the
// next power of two); the rank of a key is the dense id of its group: slot_gid[s][slot] = rank, ngroups[s] = their number
// (capped at GH_MAX_GROUPS, overflow flagged), gkeys[s][rank] = key, gw[s][rank] = (w_ref, w_alt, w_het, 0) of the group.
template <typename KeyT>
__global__ void __launch_bounds__(1024) k_group_rank(const unsigned long long *__restrict__ hash, const double *__restrict__ wtable, int32_t code_bits,
                                                     uint16_t *__restrict__ slot_gid, int32_t *__restrict__ ngroups, KeyT *__restrict__ gkeys,
                                                     double4 *__restrict__ gw, int *__restrict__ overflow) {
    __shared__ unsigned long long sm[GH_SLOTS];
    __shared__ int s_warp[33];
    const int s = blockIdx.x;
    const unsigned long long *tab = hash + size_t(s) * GH_SLOTS;
    // compaction: thread -> 4 consecutive slots
    unsigned long long v[4];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j] = tab[4 * threadIdx.x + j];
        cnt += v[j] != 0ull;
    }
    int n;
    int at = block_excl_scan(cnt, &n, s_warp);
    int P = 1;
    while (P < n) P <<= 1;
    for (int j = threadIdx.x; j < P; j += blockDim.x) sm[j] = ~0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (v[j] != 0ull) sm[at++] = ((v[j] - 1ull) << 12) | (unsigned long long)(4 * threadIdx.x + j);      // keys have at most 50 bits
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int x = i ^ j;
                if (x > i) {
                    const unsigned long long a = sm[i], c = sm[x];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) { sm[i] = c; sm[x] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int ng = min(n, GH_MAX_GROUPS);
    uint16_t *out = slot_gid + size_t(s) * GH_SLOTS;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const unsigned long long e = sm[p];
        out[e & 4095ull] = uint16_t(min(p, GH_MAX_GROUPS - 1));
        if (p < ng) {
            const KeyT key = KeyT(e >> 12);
            gkeys[size_t(s) * GH_MAX_GROUPS + p] = key;
            double4 w;
            w.x = __ldg(wtable + gs_code(key, 0, code_bits));
            w.y = __ldg(wtable + gs_code(key, 1, code_bits));
            w.z = __ldg(wtable + gs_code(key, 2, code_bits));
            w.w = 0.0;
            gw[size_t(s) * GH_MAX_GROUPS + p] = w;
        }
    }
    if (threadIdx.x == 0) {
        ngroups[s] = ng;
        if (n > GH_MAX_GROUPS) overflow[s] = 1;
    }
}

// One CTA per tile: dense id of every pair (lookup in its sample's table) -> gid[], and the tile's id counts -> tile_hist[t][0..ngroups)
// (rows of GH_MAX_GROUPS counters; only the first ngroups[s] are written and read).
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) k_group_ids(const KeyT *__restrict__ key, const int2 *__restrict__ range,
                                                         const int32_t *__restrict__ tile_sample, const unsigned long long *__restrict__ hash,
                                                         const uint16_t *__restrict__ slot_gid, const int32_t *__restrict__ ngroups,
                                                         uint16_t *__restrict__ gid, uint32_t *__restrict__ tile_hist) {
    __shared__ uint32_t h[GH_MAX_GROUPS];
    const int t = blockIdx.x;
    const int2 rg = range[t];
    const int begin = rg.x, end = rg.y;
    const int s = tile_sample[t];
    const int ng = ngroups[s];
    if (begin >= end) {                                             // an empty tile (the layout is an upper bound): its counts are zero
        for (int d = threadIdx.x; d < ng; d += RS_THREADS) tile_hist[size_t(t) * GH_MAX_GROUPS + d] = 0u;
        return;
    }
    for (int d = threadIdx.x; d < ng; d += RS_THREADS) h[d] = 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long *tab = hash + size_t(s) * GH_SLOTS;
    const uint16_t *sg = slot_gid + size_t(s) * GH_SLOTS;
    KeyT k[RS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        k[e] = i < end ? key[i] : KeyT(0);
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        if (i < end) {
            const uint32_t g = min(uint32_t(__ldg(sg + gh_find(tab, (unsigned long long)(k[e])))), uint32_t(max(ng, 1) - 1));
            gid[i] = uint16_t(g);
            atomicAdd(h + g, 1u);
        }
    }
    __syncthreads();
    uint32_t *out = tile_hist + size_t(t) * GH_MAX_GROUPS;
    for (int d = threadIdx.x; d < ng; d += RS_THREADS) out[d] = h[d];
}

// One CTA per sample: tile_hist[t][g] (pairs of id g in tile t) -> pairs of id g in EARLIER tiles of the sample, and
// goff[s][g] = first position of group g inside the sample's grouped range (goff[s][ngroups] = its pairs).
__global__ void __launch_bounds__(1024) k_group_scan(uint32_t *__restrict__ tile_hist, const int32_t *__restrict__ tile_first,
                                                     const int32_t *__restrict__ ngroups, int32_t *__restrict__ goff) {
    __shared__ int s_warp[33];
    const int s = blockIdx.x;
    const int ng = ngroups[s];
    const int t0 = tile_first[s], t1 = tile_first[s + 1];
    constexpr int PITCH = GH_MAX_GROUPS, U = 8;
    uint32_t tot[2] = {0u, 0u};                                   // thread -> ids 2 tid, 2 tid + 1
    const int d0 = 2 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (d0 + j < ng) {
            uint32_t run = 0u;
            for (int t = t0; t < t1; t += U) {
                uint32_t c[U];
#pragma unroll
                for (int u = 0; u < U; ++u) c[u] = t + u < t1 ? tile_hist[size_t(t + u) * PITCH + d0 + j] : 0u;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (t + u < t1) tile_hist[size_t(t + u) * PITCH + d0 + j] = run;
                    run += c[u];
                }
            }
            tot[j] = run;
        }
    }
    int all;
    const int ex = block_excl_scan(int(tot[0] + tot[1]), &all, s_warp);
    int32_t *out = goff + size_t(s) * (GH_MAX_GROUPS + 1);
    if (d0 < ng) out[d0] = ex;
    if (d0 + 1 < ng) out[d0 + 1] = ex + int(tot[0]);
    if (threadIdx.x == 0) out[ng] = all;
}

// One CTA per tile: the stable partition of the tile's pairs by dense id:
//   offsets  where the tile's pairs of id g go inside the sample's range = goff[s][g] + (pairs of id g in earlier tiles), both
//            from k_group_scan;
//   ranks    warp w owns pairs [256 w, 256 w + 256) of the tile and takes them 32 at a time in order, so ranks inside an id
//            follow the pair order: (pairs of that id in earlier warps) + (earlier rounds of this warp) + (lower lanes of this
//            round, match.any);
//   scatter  the panel row of the pair (and, when the caller wants the pairs back, its marker index) to its place.  Nothing
//            else moves: keys and weights live in the per-sample group table.
// Dynamic shared memory: GH_MAX_GROUPS * 20 bytes.
__global__ void __launch_bounds__(RS_THREADS, 3) k_group_place(const uint16_t *__restrict__ gid_in,
                                                              const int32_t *__restrict__ pair_db_in, const int32_t *__restrict__ pair_s_in,
                                                              int32_t *__restrict__ pair_db_out, int32_t *__restrict__ pair_s_out,
                                                              const int32_t *__restrict__ mstart, const int32_t *__restrict__ tile_sample,
                                                              const int32_t *__restrict__ tile_first, const int2 *__restrict__ range,
                                                              const uint32_t *__restrict__ tile_hist, const int32_t *__restrict__ ngroups,
                                                              const int32_t *__restrict__ goff, unsigned int *__restrict__ work_counter) {
    extern __shared__ uint32_t rs_sm[];
    if (blockIdx.x == 0 && threadIdx.x == 0) *work_counter = 0u;      // the scoring kernel's ticket counter (it runs next on this stream)
    constexpr int NW = RS_THREADS / 32;
    constexpr int PITCH = GH_MAX_GROUPS;
    uint16_t *whist = reinterpret_cast<uint16_t *>(rs_sm);          // [NW][PITCH]
    uint32_t *off = rs_sm + NW * PITCH / 2;                         // [PITCH]
    const int t = blockIdx.x;
    const int2 rg = range[t];
    const int begin = rg.x, end = rg.y;
    const int s = tile_sample[t];
    if (begin >= end) return;                                       // an empty tile (the whole CTA)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t d[RS_PER_THREAD];
    int32_t pdb[RS_PER_THREAD], ps[RS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const int i = begin + warp * (32 * RS_PER_THREAD) + e * 32 + lane;
        const bool on = i < end;
        d[e] = on ? uint32_t(gid_in[i]) : 0xffffffffu;
        pdb[e] = on ? pair_db_in[i] : 0;
        ps[e] = on && pair_s_out ? pair_s_in[i] : 0;
    }
    const int ng = ngroups[s];
    const int base = mstart[s];
    for (int j = threadIdx.x; j < NW * PITCH / 2; j += RS_THREADS) rs_sm[j] = 0u;
    for (int g = threadIdx.x; g < ng; g += RS_THREADS)
        off[g] = uint32_t(__ldg(goff + size_t(s) * (GH_MAX_GROUPS + 1) + g)) + __ldg(tile_hist + size_t(t) * PITCH + g);
    __syncthreads();
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint16_t *mine_h = whist + size_t(warp) * PITCH;
    uint32_t rank[RS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        const bool on = d[e] != 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d[e]);
        uint32_t prior = 0u;
        if (on) prior = mine_h[d[e]];
        __syncwarp();
        if (on && lane == __ffs(peers) - 1) mine_h[d[e]] = uint16_t(prior + __popc(peers));
        __syncwarp();
        rank[e] = prior + uint32_t(__popc(peers & lt_mask));
    }
    __syncthreads();
    for (int g = threadIdx.x; g < ng; g += RS_THREADS) {
        uint32_t run = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t c = whist[size_t(w) * PITCH + g];
            whist[size_t(w) * PITCH + g] = uint16_t(run);
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < RS_PER_THREAD; ++e) {
        if (d[e] != 0xffffffffu) {
            const int o = base + int(off[d[e]]) + int(mine_h[d[e]]) + int(rank[e]);
            pair_db_out[o] = pdb[e];
            if (pair_s_out) pair_s_out[o] = ps[e];
        }
    }
}

// Work order of the scoring kernel: the segments of the batch by descending cost class (16 classes by the ratio of the estimate
// to the cost of a segment without weight changes), inside a class by position in the sample (segment j of every sample, then
// j + 1, ...: SO_LEVELS levels), each entry with everything a team needs to start: (segment, first pair, rows, sample).
// Expensive segments — rare weight triples, a few rows per group — start first and the seven teams of a CTA, which run a round
// in lockstep, get neighbours of equal cost; the order inside a class keeps what the plain j-major order had for called
// genotypes, where a sample's rows are in position order inside three big groups: the CTAs gather from the same stretch of the
// panel at the same time (measured: 0.205 ms against 0.217 ms with the segments of a class in sample-major order).  A counting
// sort over (class, level) buckets: k_group_marks counts and ranks (global atomics: the order inside a bucket is whatever they
// give), k_order_place scans the counts and writes the entries.  The order decides who scores a segment when, never a result
// (every segment has its own partial sums).
constexpr int SEG_CLASSES = 16;
constexpr int SO_LEVELS = 128;
constexpr int SO_BUCKETS = SEG_CLASSES * SO_LEVELS;
__device__ __forceinline__ int seg_cost_class(int cost, int base) {       // 0 = most expensive
    const int r = (cost * 16) / max(base, 1);                              // cost / base in sixteenths
    const int lim[SEG_CLASSES - 1] = {384, 256, 192, 160, 128, 96, 80, 64, 48, 40, 32, 28, 24, 20, 18};
    int c = 0;
#pragma unroll
    for (int k = 0; k < SEG_CLASSES - 1; ++k) c += r < lim[k] ? 1 : 0;
    return c;
}

// One CTA per sample, from the group table alone: per block of 16 grouped rows of the sample (blocks counted from its first
// pair; a segment of `chunk` rows is chunk/16 blocks) one word
//     ref mask | alt mask << 16 | het mask << 32 | (group of the block's first row) << 48
// mask bit k set <=> the weight of that class at row 16 b + k differs from the row before it.  blk_chg must be zeroed before.
// Also the cost estimate of every segment for the scoring kernel's work order: blocks of 16 rows and class weight changes, in
// units of SEG_COST_*, and from it the segment's bucket and its rank inside the bucket (seg_cost and bucket_cnt zeroed before).
constexpr int SEG_COST_BLOCK = 8;         // one block of 16 rows without a weight change (~0.85 us of a team)
constexpr int SEG_COST_CHANGE = 8;        // one class weight change: a piece of a block re-added under a mask + a counter read-out
// the body of k_group_marks for sample s; s_off[0..ng] (shared memory, complete and synchronised): the group starts of the sample
template <typename KeyT>
__device__ __forceinline__ void group_marks_body(int s, const int32_t *s_off, int ng, const KeyT *__restrict__ gkeys,
                                                 const int32_t *__restrict__ mstart, const int32_t *__restrict__ seg_off, int32_t chunk,
                                                 int32_t code_bits, unsigned long long *__restrict__ blk_chg, int32_t *__restrict__ seg_cost,
                                                 int32_t jcap, int32_t *__restrict__ bucket_cnt, int2 *__restrict__ seg_br) {
    const int m = mstart[s + 1] - mstart[s];
    const int seg0 = seg_off[s];
    const int nseg_s = seg_off[s + 1] - seg0;
    for (int j = threadIdx.x; j < nseg_s; j += blockDim.x) {
        const int rows = min(chunk, m - j * chunk);
        atomicAdd(seg_cost + seg0 + j, SEG_COST_BLOCK * ((rows + 15) / 16));
    }
    if (m > 0 && ng > 0) {
        unsigned long long *out = blk_chg + size_t(seg0) * size_t(chunk / 16);
        const KeyT *gk = gkeys + size_t(s) * GH_MAX_GROUPS;
        const int b = code_bits;
        for (int g = 1 + threadIdx.x; g < ng; g += blockDim.x) {       // the start of group g: which classes change there
            const int r = s_off[g];
            if (r >= m || s_off[g + 1] == r) continue;                 // (an id without pairs cannot occur; be safe)
            const KeyT k1 = gk[g], k0 = gk[g - 1];
            const int c1 = int(k1 >> (3 * b)) & 3, c0 = int(k0 >> (3 * b)) & 3;
            unsigned long long bits = 0ull;
            if (c1 != c0) {
                bits = 1ull | (1ull << 16) | (1ull << 32);
            } else {
                const KeyT dd = k1 ^ k0;
                const uint32_t fm = (1u << b) - 1u;
#pragma unroll
                for (int w = 0; w < 3; ++w)
                    if ((uint32_t(dd >> gs_field_shift(c1, w, b)) & fm) != 0u) bits |= 1ull << (16 * w);
            }
            atomicOr(out + (r >> 4), bits << (r & 15));
            atomicAdd(seg_cost + seg0 + r / chunk, SEG_COST_CHANGE * __popcll(bits));
        }
        const int nb = (m + 15) / 16;
        for (int blk = threadIdx.x; blk < nb; blk += blockDim.x) {     // group of the block's first row: the last start <= 16 blk
            int lo = 0, hi = ng;
            const int r = 16 * blk;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= r) lo = mid; else hi = mid;
            }
            atomicOr(out + blk, (unsigned long long)(lo) << 48);
        }
    }
    // once the costs are complete: every segment's bucket of the work order and its rank inside the bucket
    __threadfence();
    __syncthreads();
    const int base = SEG_COST_BLOCK * (chunk / 16);
    const int jdiv = (jcap + SO_LEVELS - 1) / SO_LEVELS;
    for (int j = threadIdx.x; j < nseg_s; j += blockDim.x) {
        const int bkt = seg_cost_class(__ldcg(seg_cost + seg0 + j), base) * SO_LEVELS + min(SO_LEVELS - 1, j / jdiv);
        seg_br[seg0 + j] = make_int2(bkt, atomicAdd(bucket_cnt + bkt, 1));
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(1024) k_group_marks(const int32_t *__restrict__ goff, const KeyT *__restrict__ gkeys, const int32_t *__restrict__ ngroups,
                                                      const int32_t *__restrict__ mstart, const int32_t *__restrict__ seg_off, int32_t chunk,
                                                      int32_t code_bits, unsigned long long *__restrict__ blk_chg, int32_t *__restrict__ seg_cost,
                                                      int32_t jcap, int32_t *__restrict__ bucket_cnt, int2 *__restrict__ seg_br) {
    __shared__ int32_t s_off[GH_MAX_GROUPS + 1];
    const int s = blockIdx.x;
    const int ng = ngroups[s];
    for (int g = threadIdx.x; g <= ng; g += blockDim.x) s_off[g] = goff[size_t(s) * (GH_MAX_GROUPS + 1) + g];
    __syncthreads();
    group_marks_body<KeyT>(s, s_off, ng, gkeys, mstart, seg_off, chunk, code_bits, blk_chg, seg_cost, jcap, bucket_cnt, seg_br);
}

// ---- samples of up to GP_MAX_PAIRS matched pairs: ids, offsets, placement and block words in ONE kernel, one CTA per sample ----
// What k_tile_ranges + k_group_ids + k_group_scan + k_group_place + k_group_marks do over tiles of 2048 pairs (five launches,
// 0.07 ms of mostly latency per 2.9 M pairs) for a sample whose counters fit shared memory: warp w owns the w-th 32nd of the
// sample's pairs (in pair order);
//   pass 1   dense id of every pair (lookup in the sample's key table, four independent probes in flight per lane) -> gid[],
//            counts per (warp, id) in shared memory (u16 [32][2048]; lanes with equal ids elect one: match.any);
//   scan     per id: exclusive prefix of the counts over the warps (in place) and the id's total; block scan of the totals ->
//            s_off[id] = first position of the group inside the sample (also written to goff);
//   pass 2   every warp walks its pairs again in order: position = s_off[id] + (pairs of that id in earlier warps) + (earlier
//            rounds of this warp) + (lower lanes of this round): a stable partition, deterministic; the panel row (and marker
//            index) of the pair goes to its place;
//   marks    group_marks_body: block words, segment costs, buckets of the work order, from s_off in shared memory.
// Dynamic shared memory: GP_SMEM bytes.  Counts and prefixes are 16-bit: a sample must not hold more than GP_MAX_PAIRS pairs
// (the caller falls back to the tiled kernels for batches with larger samples).
constexpr int GP_WARPS = 32;
constexpr int GP_MAX_PAIRS = 65535;
constexpr size_t GP_SMEM = size_t(GP_WARPS) * GH_MAX_GROUPS * 2 + size_t(GH_MAX_GROUPS + 4) * 4;
template <typename KeyT>
__global__ void __launch_bounds__(32 * GP_WARPS) k_group_sample(
        const KeyT *__restrict__ key, const unsigned long long *__restrict__ hash, const uint16_t *__restrict__ slot_gid,
        const int32_t *__restrict__ ngroups, uint16_t *__restrict__ gid, const int32_t *__restrict__ pair_db_in,
        const int32_t *__restrict__ pair_s_in, int32_t *__restrict__ pair_db_out, int32_t *__restrict__ pair_s_out,
        int32_t *__restrict__ goff, const KeyT *__restrict__ gkeys, const int32_t *__restrict__ mstart,
        const int32_t *__restrict__ seg_off, int32_t chunk, int32_t code_bits, unsigned long long *__restrict__ blk_chg,
        int32_t *__restrict__ seg_cost, int32_t jcap, int32_t *__restrict__ bucket_cnt, int2 *__restrict__ seg_br,
        unsigned int *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char gp_smem[];
    __shared__ int s_warp[33];
    uint16_t *whist = reinterpret_cast<uint16_t *>(gp_smem);                                  // [GP_WARPS][GH_MAX_GROUPS]
    uint32_t *whist32 = reinterpret_cast<uint32_t *>(gp_smem);                                // the same, two ids per word
    int32_t *s_off = reinterpret_cast<int32_t *>(gp_smem + size_t(GP_WARPS) * GH_MAX_GROUPS * 2);   // [GH_MAX_GROUPS + 1]
    if (blockIdx.x == 0 && threadIdx.x == 0) *work_counter = 0u;      // the scoring kernel's ticket counter (it runs next on this stream)
    const int s = blockIdx.x;
    const int ng = ngroups[s];
    const int b0 = mstart[s], m = min(mstart[s + 1] - b0, GP_MAX_PAIRS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int used_words = (ng + 1) / 2;                              // words of a warp's row that hold ids of this sample
    for (int j = threadIdx.x; j < GP_WARPS * used_words; j += blockDim.x) whist32[size_t(j / used_words) * (GH_MAX_GROUPS / 2) + j % used_words] = 0u;
    __syncthreads();
    const int per = ((m + GP_WARPS - 1) / GP_WARPS + 31) & ~31;       // pairs per warp, whole rounds of 32
    const int r0 = min(m, warp * per), r1 = min(m, r0 + per);
    uint16_t *mine = whist + size_t(warp) * GH_MAX_GROUPS;
    const unsigned long long *tab = hash + size_t(s) * GH_SLOTS;
    const uint16_t *sg = slot_gid + size_t(s) * GH_SLOTS;
    const uint32_t last_id = uint32_t(max(ng, 1) - 1);
    constexpr int U = 8;
    for (int base = r0; base < r1; base += 32 * U) {
        unsigned long long want[U], cur[U];
        uint32_t h[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + 32 * u + lane;
            want[u] = i < r1 ? (unsigned long long)(key[b0 + i]) + 1ull : 0ull;
            h[u] = gh_hash(want[u] - 1ull);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = want[u] ? __ldg(tab + h[u]) : 0ull;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            for (int probe = 1; probe < GH_SLOTS && cur[u] != want[u] && cur[u] != 0ull; ++probe) {
                h[u] = (h[u] + 1u) & uint32_t(GH_SLOTS - 1);
                cur[u] = __ldg(tab + h[u]);
            }
        }
        uint32_t g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) g[u] = want[u] ? min(uint32_t(__ldg(sg + h[u])), last_id) : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (base + 32 * u >= r1) break;                          // uniform: the whole round is past the warp's range
            const int i = base + 32 * u + lane;
            const bool on = i < r1;
            if (on) gid[b0 + i] = uint16_t(g[u]);
            const uint32_t peers = __match_any_sync(0xffffffffu, g[u]);
            if (on && lane == __ffs(peers) - 1) mine[g[u]] = uint16_t(mine[g[u]] + __popc(peers));
            __syncwarp();
        }
    }
    __syncthreads();
    // per id: prefix of the counts over the warps, total; thread -> ids 2 tid, 2 tid + 1 (one 32-bit word of every warp's row)
    uint32_t tot0 = 0u, tot1 = 0u;
    if (2 * int(threadIdx.x) < ng) {
#pragma unroll 8
        for (int w = 0; w < GP_WARPS; ++w) {
            const uint32_t c = whist32[size_t(w) * (GH_MAX_GROUPS / 2) + threadIdx.x];
            whist32[size_t(w) * (GH_MAX_GROUPS / 2) + threadIdx.x] = tot0 | (tot1 << 16);
            tot0 += c & 0xffffu;
            tot1 += c >> 16;
        }
    }
    int all;
    const int ex = block_excl_scan(int(tot0 + tot1), &all, s_warp);
    int32_t *out = goff + size_t(s) * (GH_MAX_GROUPS + 1);
    const int d0 = 2 * int(threadIdx.x);
    if (d0 < ng) { s_off[d0] = ex; out[d0] = ex; }
    if (d0 + 1 < ng) { s_off[d0 + 1] = ex + int(tot0); out[d0 + 1] = ex + int(tot0); }
    if (threadIdx.x == 0) { s_off[ng] = all; out[ng] = all; }
    __syncthreads();
    // placement (the loads of U rounds are issued before the first round is ranked: a warp has nothing else to hide them behind)
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int base = r0; base < r1; base += 32 * U) {
        uint32_t g[U];
        int32_t pdb[U], ps[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + 32 * u + lane;
            const bool on = i < r1;
            g[u] = on ? uint32_t(gid[b0 + i]) : 0xffffffffu;
            pdb[u] = on ? pair_db_in[b0 + i] : 0;
            ps[u] = on && pair_s_out ? pair_s_in[b0 + i] : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (base + 32 * u >= r1) break;                          // uniform
            const bool on = g[u] != 0xffffffffu;
            const uint32_t peers = __match_any_sync(0xffffffffu, g[u]);
            uint32_t prior = 0u;
            if (on) prior = mine[g[u]];
            __syncwarp();
            if (on && lane == __ffs(peers) - 1) mine[g[u]] = uint16_t(prior + __popc(peers));
            __syncwarp();
            if (on) {
                const int o = b0 + s_off[g[u]] + int(prior) + __popc(peers & lt_mask);
                pair_db_out[o] = pdb[u];
                if (pair_s_out) pair_s_out[o] = ps[u];
            }
        }
    }
    __syncthreads();
    group_marks_body<KeyT>(s, s_off, ng, gkeys, mstart, seg_off, chunk, code_bits, blk_chg, seg_cost, jcap, bucket_cnt, seg_br);
}

// One CTA per sample: exclusive scan of the bucket counts (every CTA for itself: 8 KB), then the entries of the sample's segments.
__global__ void __launch_bounds__(256) k_order_place(const int32_t *__restrict__ bucket_cnt, const int2 *__restrict__ seg_br, const int32_t *__restrict__ seg_off,
                                                    const int32_t *__restrict__ mstart, int32_t chunk, int4 *__restrict__ order) {
    __shared__ int scan[SO_BUCKETS];
    __shared__ int warp_sum[8];
    const int s = blockIdx.x, t = threadIdx.x;
    const int seg0 = seg_off[s], nseg_s = seg_off[s + 1] - seg0;
    if (nseg_s <= 0) return;
    constexpr int PT = SO_BUCKETS / 256;
    int v[PT], sum = 0;
#pragma unroll
    for (int k = 0; k < PT; ++k) { v[k] = __ldcg(bucket_cnt + PT * t + k); sum += v[k]; }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, d);
        if ((t & 31) >= d) incl += u;
    }
    if ((t & 31) == 31) warp_sum[t >> 5] = incl;
    __syncthreads();
    int run = incl - sum;
    for (int w = 0; w < (t >> 5); ++w) run += warp_sum[w];
#pragma unroll
    for (int k = 0; k < PT; ++k) { scan[PT * t + k] = run; run += v[k]; }
    __syncthreads();
    const int m0 = mstart[s], m1 = mstart[s + 1];
    for (int j = t; j < nseg_s; j += blockDim.x) {
        const int2 br = seg_br[seg0 + j];
        const int begin = m0 + j * chunk;
        order[scan[br.x] + br.y] = make_int4(seg0 + j, begin, min(m1, begin + chunk) - begin, s);
    }
}

}  // namespace snpm
