// k_score_grouped2 — the counting kernel (A2/A3 in grouped order) as a persistent grid over device-grouped pairs.
//
// Same arithmetic as k_score_grouped (grouped.cuh): a thread owns one 32-accession word column of a segment, adds the class
// planes of 16 rows at a time into bit-sliced counters, reads a class counter out only where that class's weight changes
// (per-block change masks, here precomputed by k_group_masks), folds counts into fp64 with one fma per accession, keeps
// weight-1.0 classes in an exact integer counter: score = I + F (see grouped.cuh for the exactness argument).
// What is different:
//   * input is what the device-side grouping produces (group_sort.cuh): panel rows in grouped order, one word per 16-row
//     block (change masks + the group its first row belongs to) and the weights of every group (a few KB per sample);
//   * ONE persistent CTA per SM, 7 teams of 36 threads in 8 warps (252 of 256 lanes work: the 128-thread CTAs of round 1
//     used 108 of 128) and 7 instead of 6 segments in flight per SM; wide panels (rows of more than 36 words) are cut into
//     warp-aligned slices of 32 words, 8 teams = 8 slices of the same segment;
//   * the CTA works in rounds: it draws `teams` consecutive (segment, word slice) items from a global counter, longest
//     segments first; the teams score them block by block in lockstep (one full-warp barrier per 16-row block, a no-op
//     while the warp is converged) and meet at ONE __syncthreads per round.  Lockstep is deliberate: a 36-thread team
//     spans two warps, and two teams that share a warp issue ONE instruction stream only while they run the same code
//     (measured: independent teams executed 160 M warp instructions for the work the lockstep kernel does in 113 M);
//   * the row numbers, keys and change masks of the NEXT round are fetched with cp.async while the current one is scored
//     (double-buffered staging), so no team waits for a dependent global load between segments;
//   * class counters and totals have 9 planes (segments of at most 496 rows: no counter can overflow inside one).
#pragma once
#include "common.cuh"
#include "grouped.cuh"
#include "group_sort.cuh"

namespace snpm {

constexpr int G2_WX = 36;                 // words per team (one 1135-accession row)
constexpr int G2_MAX_TEAMS = 8;
constexpr int G2_THREADS = 256;           // 7 teams x 36 = 252 threads: 8 warps, two per scheduler, so a thread may hold 255 registers
                                          // (9 warps would put three on one scheduler's register file: 168 registers, spills)
constexpr int G2_MAX_CHUNK = 496;
constexpr int G2_CP = 9;                  // planes of a class counter
constexpr int G2_TP = 9;                  // planes of the per-segment totals (I, ninfo)

struct Group2Args {
    const uint64_t *packed;
    int32_t stride;
    const int32_t *pair_db;               // [m] matched local rows, grouped order
    const unsigned long long *blk_chg;    // per 16-row block: change masks + group of its first row (k_group_marks): [segment][chunk / 16]
    const double *gw;                     // [S][GH_MAX_GROUPS][4] weights (ref, alt, het, 0) of every group of every sample
    const int32_t *seg_off;               // [S+1]
    const int32_t *mstart;                // [S+1]
    int32_t S;
    int32_t chunk;
    double *part_score;                   // [nseg, 32, stride]  F of the segment; accession 32 w + b at [b][w]
    int32_t *part_int;                    // [nseg, 32, stride]  low half I, high half ninfo
    int32_t a_pad;
    int32_t wx;                           // words per team slice
    int32_t teams;                        // teams per CTA
    int32_t n_slices;                     // word slices of a row
    const int4 *order;                    // work order: (segment, first pair, rows, sample), most expensive first (k_order_place)
    unsigned int *work_counter;           // zeroed before the launch
};

__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}

// one work item: a segment (chunk rows of one sample) x one word slice
struct G2Item {
    int32_t seg;        // -1: no more work
    int32_t begin;      // first pair
    int32_t n_rows;
    int32_t slice;
    int32_t smp;
};

__host__ __device__ __forceinline__ size_t g2_stage_bytes(int chunk) {
    return (size_t(chunk) * 4 + size_t(chunk / 16 + 1) * 8 + 15) & ~size_t(15);      // row numbers | block words
}
__host__ __device__ __forceinline__ size_t g2_team_smem(int wx, int chunk, int ring) {
    return size_t(ring) * wx * 8 + 2 * g2_stage_bytes(chunk);                // ring | 2 staging buffers
}

// the four bytes of a word as doubles (I2F.F64.U8 takes a byte of a register; written as PTX because the compiler's own
// lowering of (t >> 8j) & 0xff spends a shift and a mask per byte before the conversion)
__device__ __forceinline__ void bytes_to_f64(uint32_t t, double &d0, double &d1, double &d2, double &d3) {
    asm("{\n\t.reg .b8 b0, b1, b2, b3;\n\tmov.b32 {b0, b1, b2, b3}, %4;\n\tcvt.rn.f64.u8 %0, b0;\n\tcvt.rn.f64.u8 %1, b1;\n\t"
        "cvt.rn.f64.u8 %2, b2;\n\tcvt.rn.f64.u8 %3, b3;\n\t}"
        : "=d"(d0), "=d"(d1), "=d"(d2), "=d"(d3)
        : "r"(t));
}

// row k of a block if bit k of `rm` is set, else 0: one predicate per row (R2P) and a select, instead of shift, shift, and
template <int K>
__device__ __forceinline__ uint32_t row_if_bit(uint32_t pl, uint32_t rm) {
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\tselp.b32 %0, %1, 0, p;\n\t}"
        : "=r"(r)
        : "r"(pl), "r"(rm), "n"(1u << K));
    return r;
}
template <int K>
struct MaskRows {
    static __device__ __forceinline__ void go(uint32_t (&m)[GR_BLOCK], const uint32_t (&pl)[GR_BLOCK], uint32_t rm) {
        m[K] = row_if_bit<K>(pl[K], rm);
        MaskRows<K - 1>::go(m, pl, rm);
    }
};
template <>
struct MaskRows<-1> {
    static __device__ __forceinline__ void go(uint32_t (&)[GR_BLOCK], const uint32_t (&)[GR_BLOCK], uint32_t) {}
};

// fold the counts of a class counter: F[lane] += w * count[lane]
__device__ __forceinline__ void fold_counts9(const BitCounter<G2_CP> &c, double w, double (&F)[32]) {
    uint32_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = c.p[k];
    transpose_planes8(t);
    if (c.p[8] == 0u) {                           // fewer than 256 rows since the last read-out: the common case
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double d0, d1, d2, d3;
            bytes_to_f64(t[i], d0, d1, d2, d3);
            F[i] = fma(w, d0, F[i]);
            F[8 + i] = fma(w, d1, F[8 + i]);
            F[16 + i] = fma(w, d2, F[16 + i]);
            F[24 + i] = fma(w, d3, F[24 + i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = 8 * j + i;
                const int cnt = int((t[i] >> (8 * j)) & 0xffu) | int(((c.p[8] >> b) & 1u) << 8);
                F[b] = fma(w, double(cnt), F[b]);
            }
        }
    }
}

__device__ __forceinline__ void counter_values9(const BitCounter<G2_TP> &c, int32_t (&v)[32]) {
    uint32_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = c.p[k];
    transpose_planes8(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = 8 * j + i;
            v[b] = int32_t((t[i] >> (8 * j)) & 0xffu) | int32_t(((c.p[8] >> b) & 1u) << 8);
        }
    }
}

// WX > 0: words per team known at compile time; WX == 0: a.wx.  Thread pairs (2i, 2i+1) share 16-byte copies of two columns.
// (Measured and dropped: pulling the rows of block b + 6 / b + 12 into L2 with prefetch.global.L2 while block b's copies are
// queued — to have more bytes in flight at the DRAM level than the shared-memory ring holds — made the kernel slower: 0.272 ms
// without, 0.328 ms at distance 6, 0.354 ms at 12.)
// RING: rows of a team's gather ring (a multiple of 16): RING / 16 - 1 blocks are in flight while one is scored.  (Measured:
// 96 rows instead of 64 changed nothing for called genotypes (0.205 ms) and cost 4-6 % with PL weights and on the 20 000-accession
// panel; 10 teams in a 384-thread CTA at 168 registers — three warps per scheduler, 224 bytes of spills — ran at 0.350 ms against
// 0.307 in the same session.  Drawing the next round's ticket one round ahead instead of two — published by thread 0 through
// shared memory and picked up mid-round, to shorten the tail (the slowest SM is active 26 % longer than the average one) —
// was slower everywhere: 0.2764 vs 0.2702 ms, 0.981 vs 0.924 ms on the 20 000-accession panel.)
template <bool SKIP_HETS, int WX, int RING>
__global__ void __launch_bounds__(G2_THREADS, 1) k_score_grouped2(const Group2Args a) {
    constexpr int INFLIGHT = RING / GR_BLOCK;
    extern __shared__ __align__(16) unsigned char g2_smem[];
    __shared__ int s_base[2];                     // first item of the round in staging buffer 0 / 1
    const int wx = WX ? WX : a.wx;
    const int q = threadIdx.x / wx, w = threadIdx.x - q * wx;
    const bool in_team = q < a.teams;
    const size_t team_bytes = g2_team_smem(wx, a.chunk, RING);
    const size_t stage_bytes = g2_stage_bytes(a.chunk);
    unsigned char *team = g2_smem + size_t(in_team ? q : 0) * team_bytes;
    uint64_t *ring = reinterpret_cast<uint64_t *>(team);
    unsigned char *stage0 = team + size_t(RING) * wx * 8;
    const int n_items = __ldg(a.seg_off + a.S) * a.n_slices;      // segments of the batch x word slices
    // item i -> (entry i / n_slices of the work order, slice i % n_slices)
    auto decode = [&](int i) {
        G2Item it;
        it.seg = -1; it.begin = 0; it.n_rows = 0; it.slice = 0; it.smp = 0;
        if (in_team && i < n_items) {
            const int rank = i / a.n_slices;
            const int4 o = __ldg(a.order + rank);
            it.seg = o.x;
            it.begin = o.y;
            it.n_rows = o.z;
            it.slice = i - rank * a.n_slices;
            it.smp = o.w;
        }
        return it;
    };
    // queue the copies of an item's row numbers, keys and change masks into staging buffer `buf` (one commit group)
    auto prefetch = [&](const G2Item &it, int buf) {
        if (it.seg >= 0) {
            unsigned char *st = stage0 + size_t(buf) * stage_bytes;
            const uint32_t s_row = smem_u32(st), s_chg = s_row + uint32_t(a.chunk) * 4u;
            for (int r = w; r < it.n_rows; r += wx) cp_async4(s_row + uint32_t(r) * 4u, a.pair_db + it.begin + r);
            const int nb = (it.n_rows + GR_BLOCK - 1) / GR_BLOCK;
            const unsigned long long *src = a.blk_chg + size_t(it.seg) * size_t(a.chunk / GR_BLOCK);
            for (int b = w; b < nb; b += wx) cp_async8(s_chg + uint32_t(b) * 8u, src + b);
        }
        cp_async_commit();
    };

    // start-up: round 0 staged synchronously.  The first three rounds of a CTA are fixed — a ticket is drawn two rounds before
    // it is used, so three rounds are committed when the kernel starts whatever the scheme: round blockIdx.x of the work order
    // (most expensive first) and two of the 2 G cheapest rounds at its end (R - 1 - blockIdx.x, R - 1 - G - blockIdx.x), so
    // that whoever starts on the most expensive segments (184 us for one round against 215 us of average busy time) is committed
    // to ~30 us more, not to whatever its first two tickets bring.  Tickets cover rounds [G, R - 2 G) in order; the kernel ends on
    // ordinary cheap rounds.  (Three tickets per CTA at the start: the slowest CTA finished at 251 us; rounds b and 2 G - 1 - b
    // fixed: 239 us, because the cost estimate cannot see how sparse the class planes are.)
    const int G = int(gridDim.x), T = a.teams;
    const int R = (n_items + T - 1) / T;          // rounds of the launch
    const int dyn_end = max(G, R - 2 * G);        // tickets: rounds [G, dyn_end)
    const int second = R - 1 - int(blockIdx.x), third = R - 1 - G - int(blockIdx.x);
    if (threadIdx.x == 0) {
        s_base[0] = int(blockIdx.x) * T;
        s_base[1] = second >= max(G, R - G) ? second * T : n_items;
    }
    __syncthreads();
    if (s_base[0] >= n_items) return;
    G2Item cur = decode(s_base[0] + q);
    prefetch(cur, 0);
    cp_async_wait<0>();
    __syncthreads();
    int buf = 0;
    bool first_round = true;

    const int64_t stride = a.stride;
    const int odd = w & 1;
    const uint32_t ring_pitch = uint32_t(wx) * 8u;
    const uint32_t stride_b = uint32_t(a.stride) * 8u;
    const int nb_round = a.chunk / GR_BLOCK;      // every team runs this many block steps per round (lockstep)

    while (true) {
        const int next_base = s_base[buf ^ 1];
        int ticket = 0;
        if (threadIdx.x == 0 && next_base < n_items) {      // the round after next
            if (first_round) {
                ticket = third >= dyn_end && third < R - G ? third * T : n_items;
            } else {
                const int r = G + int(atomicAdd(a.work_counter, 1u));
                ticket = r < dyn_end ? r * T : n_items;
            }
        }
        first_round = false;
        const G2Item nxt = decode(next_base + q);
        prefetch(nxt, buf ^ 1);                   // lands while this round is scored (oldest commit group)
        const unsigned char *st = stage0 + size_t(buf) * stage_bytes;
        const int32_t *s_row = reinterpret_cast<const int32_t *>(st);
        const unsigned long long *s_chg = reinterpret_cast<const unsigned long long *>(st + size_t(a.chunk) * 4);
        const double *gwp = a.gw + size_t(cur.smp) * (4 * GH_MAX_GROUPS);
        const int n_rows = cur.n_rows;            // 0: no item for this team in this round (it still keeps step with the others)
        const int n_blocks = (n_rows + GR_BLOCK - 1) / GR_BLOCK;
        const int n_full = n_rows / GR_BLOCK;
        const int word = cur.slice * wx + w;
        const bool live = in_team && n_rows > 0 && word < a.stride;
        const unsigned char *col = reinterpret_cast<const unsigned char *>(a.packed + (word & ~1));
        const uint32_t my_ring = smem_u32(ring + (w & ~1));
        auto issue = [&](int b) {
            if (live) {
                const int r0 = b * GR_BLOCK + 8 * odd;
                const uint32_t slot0 = my_ring + uint32_t(r0 % RING) * ring_pitch;
                if (b < n_full) {
#pragma unroll
                    for (int k4 = 0; k4 < GR_BLOCK / 2; k4 += 4) {
                        const int4 rr = *reinterpret_cast<const int4 *>(s_row + r0 + k4);
                        cp_async16(slot0 + uint32_t(k4 + 0) * ring_pitch, col + (unsigned long long)(uint32_t(rr.x)) * stride_b);
                        cp_async16(slot0 + uint32_t(k4 + 1) * ring_pitch, col + (unsigned long long)(uint32_t(rr.y)) * stride_b);
                        cp_async16(slot0 + uint32_t(k4 + 2) * ring_pitch, col + (unsigned long long)(uint32_t(rr.z)) * stride_b);
                        cp_async16(slot0 + uint32_t(k4 + 3) * ring_pitch, col + (unsigned long long)(uint32_t(rr.w)) * stride_b);
                    }
                } else if (b < n_blocks) {
                    for (int k = 0; k < GR_BLOCK / 2 && r0 + k < n_rows; ++k)
                        cp_async16(slot0 + uint32_t(k) * ring_pitch, col + (unsigned long long)(uint32_t(s_row[r0 + k])) * stride_b);
                }
            }
            cp_async_commit();                    // always: the waits below count groups
        };
#pragma unroll
        for (int b = 0; b < INFLIGHT; ++b) issue(b);

        double F[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) F[b] = 0.0;
        BitCounter<G2_TP> c_int, c_ninfo;
        BitCounter<G2_CP> c_ref, c_alt, c_het;
        c_int.clear();
        c_ninfo.clear();
        c_ref.clear();
        c_alt.clear();
        c_het.clear();
        double w_ref = 0.0, w_alt = 0.0, w_het = 0.0;
        if (n_rows > 0) {
            const double *w0 = gwp + 4 * int(s_chg[0] >> 48);
            w_ref = __ldg(w0);
            w_alt = __ldg(w0 + 1);
            w_het = __ldg(w0 + 2);
        }
        // read a class counter out: its counts are informative sites, and matches weighted by `wt`
        auto flush_class = [&](BitCounter<G2_CP> &c, double wt) {
            if (c.any()) {
                c_ninfo.add_counter(c);
                if (wt == 1.0) c_int.add_counter(c);
                else if (wt != 0.0) fold_counts9(c, wt, F);
                c.clear();
            }
        };
        // one class: add the block's planes; where the class weight changes inside the block (bit k of `mask`: row k starts a new
        // weight), add the rows piece by piece and read the counter out in between
        auto add_class = [&](BitCounter<G2_CP> &c, double &wt, const uint32_t (&pl)[GR_BLOCK], uint32_t mask, int which, unsigned long long chg) {
            // (measured and dropped: a warp vote so that both teams of a warp take the same path — 0.2683-0.2703 ms against 0.2664)
            if (mask == 0u) {
                c.add16(pl);
                return;
            }
            int k0 = 0;
            while (true) {
                const int k1 = mask ? __ffs(mask) - 1 : GR_BLOCK;
                if (k1 > k0) {
                    const uint32_t rm = ((1u << k1) - 1u) & ~((1u << k0) - 1u);
                    uint32_t m[GR_BLOCK];
                    // measured (PL samples / 20 000-accession panel / called genotypes, kernel ms): shift-shift-and masks 0.2744 /
                    // 0.8827 / 0.1987, predicate selects 0.2664 / 0.9053 / 0.2007: the selects pay where teams share warps
                    if (WX != 32) {
                        MaskRows<GR_BLOCK - 1>::go(m, pl, rm);
                    } else {
#pragma unroll
                        for (int k = 0; k < GR_BLOCK; ++k) m[k] = pl[k] & uint32_t(int32_t(rm << (31 - k)) >> 31);
                    }
                    c.add16(m);
                }
                if (k1 >= GR_BLOCK) break;
                flush_class(c, wt);
                // the group row k1 starts in: the block's first group + the group starts up to k1 (a start at row 0 is already counted)
                const uint32_t starts = uint32_t(chg | (chg >> 16) | (chg >> 32)) & 0xffffu;
                const int g = int(chg >> 48) + __popc(starts & ((2u << k1) - 2u));
                wt = __ldg(gwp + 4 * g + which);
                mask &= mask - 1u;
                k0 = k1;
            }
        };
        auto score_block = [&](const uint32_t (&lo)[GR_BLOCK], const uint32_t (&hi)[GR_BLOCK], int b) {
            const unsigned long long chg = s_chg[b];
            uint32_t pl[GR_BLOCK];
#pragma unroll
            for (int k = 0; k < GR_BLOCK; ++k) pl[k] = ~(lo[k] | hi[k]);
            add_class(c_ref, w_ref, pl, uint32_t(chg) & 0xffffu, 0, chg);
#pragma unroll
            for (int k = 0; k < GR_BLOCK; ++k) pl[k] = lo[k] & ~hi[k];
            add_class(c_alt, w_alt, pl, uint32_t(chg >> 16) & 0xffffu, 1, chg);
            if (!SKIP_HETS) {                     // snpmatch.py:78-79: masked hets match nothing and are not informative
#pragma unroll
                for (int k = 0; k < GR_BLOCK; ++k) pl[k] = hi[k] & ~lo[k];
                add_class(c_het, w_het, pl, uint32_t(chg >> 32) & 0xffffu, 2, chg);
            }
        };

        // Every thread runs all nb_round steps (idle teams and threads past the row's last word only keep step), one full-warp
        // barrier per step: after it the neighbour's copies of block b have landed too and it has read block b-1, whose slots
        // are refilled next.
        for (int b = 0; b < nb_round; ++b) {
            cp_async_wait<INFLIGHT - 2>();
            __syncwarp();
            if (b > 0) issue(b + INFLIGHT - 1);
            if (b < n_full) {
                const uint64_t *slot = ring + size_t((b * GR_BLOCK) % RING) * wx + w;
                uint32_t lo[GR_BLOCK], hi[GR_BLOCK];
#pragma unroll
                for (int k = 0; k < GR_BLOCK; ++k) {
                    const uint64_t v = slot[size_t(k) * wx];
                    lo[k] = uint32_t(v);
                    hi[k] = uint32_t(v >> 32);
                }
                score_block(lo, hi, b);
            } else if (b < n_blocks) {            // ragged last block: rows past the end read as missing everywhere
                const int r0 = b * GR_BLOCK;
                const uint64_t *slot = ring + size_t(r0 % RING) * wx + w;
                uint32_t lo[GR_BLOCK], hi[GR_BLOCK];
#pragma unroll
                for (int k = 0; k < GR_BLOCK; ++k) {
                    uint64_t v = ~0ull;
                    if (r0 + k < n_rows) v = slot[size_t(k) * wx];
                    lo[k] = uint32_t(v);
                    hi[k] = uint32_t(v >> 32);
                }
                score_block(lo, hi, b);
            }
        }
        flush_class(c_ref, w_ref);
        flush_class(c_alt, w_alt);
        if (!SKIP_HETS) flush_class(c_het, w_het);
        if (live) {
            int32_t vi[32], vn[32];
            counter_values9(c_int, vi);
            counter_values9(c_ninfo, vn);
            const int64_t o = int64_t(cur.seg) * a.a_pad + word;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                a.part_score[o + b * stride] = F[b];
                a.part_int[o + b * stride] = vi[b] | (vn[b] << 16);
            }
        }
        // next round: its staging copies are complete (own ones: wait; everybody's: barrier), the ring is free
        cp_async_wait<0>();
        if (next_base >= n_items) break;          // the same for every thread of the CTA
        if (threadIdx.x == 0) s_base[buf] = ticket;        // nobody reads this slot during the round that ends here
        __syncthreads();                          // everyone is done with this round's staging buffer; the next base is visible
        cur = nxt;
        buf ^= 1;
    }
}

}  // namespace snpm
