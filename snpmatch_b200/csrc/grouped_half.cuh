// EXPERIMENT — compiled into the library but reached only with SNPM_GROUPED_HALF=1 in the environment; NOT YET RUN ON HARDWARE
// (written when the round's GPU budget was spent).  The half-word / row-pair variant of k_score_grouped (bit
// layouts checked by scripts/emulate_halfword_counters.py).  ptxas: 168 registers, no spills at 384 threads per CTA.
// A thread owns 16 accessions of one 32-accession word column (half h) and packs two rows per register (low 16 bits = row 2j,
// high 16 bits = row 2j+1 of a 32-row block), so BitCounter::add16 counts 32 rows per call on 32 useful bits.  A team is
// 2 * wx = 72 threads; five teams per 384-thread CTA, one CTA per SM (12 warps per SM instead of 8, 94 % of the lanes busy
// instead of 84 %).  The per-class change masks are 32 bits per block.  Same inputs and outputs as k_score_grouped.
#pragma once
#include "common.cuh"
#include "grouped.cuh"

namespace snpm {

constexpr int GH_THREADS = 384;
constexpr int GH_BLOCK = 32;                          // rows per step (16 packed planes)
constexpr int GH_RING = 128;                          // rows of the per-team ring: four blocks in flight
constexpr int GH_INFLIGHT = GH_RING / GH_BLOCK;

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// counts of the 16 accessions of a half-word thread: lane a holds the even rows, lane a + 16 the odd rows
__device__ __forceinline__ void half_counts(const BitCounter<GR_LP> &c, int32_t (&v)[16]) {
    uint32_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = c.p[k];
    transpose_planes8(t);
    const uint32_t h8 = c.p[8], h9 = c.p[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t x = t[i] & 0x00FF00FFu, y = (t[i] >> 8) & 0x00FF00FFu;
        v[i] = int32_t((x + (x >> 16)) & 0xFFFFu);
        v[8 + i] = int32_t((y + (y >> 16)) & 0xFFFFu);
    }
    if (h8 | h9) {
#pragma unroll
        for (int a = 0; a < 16; ++a)
            v[a] += int32_t((((h8 >> a) & 1u) + ((h8 >> (a + 16)) & 1u)) << 8) + int32_t((((h9 >> a) & 1u) + ((h9 >> (a + 16)) & 1u)) << 9);
    }
}

__host__ __device__ __forceinline__ size_t half_team_smem(int wx, int chunk) {
    const size_t n_blocks = size_t(chunk + GH_BLOCK - 1) / GH_BLOCK;
    return size_t(GH_RING) * wx * 8 + size_t(chunk) * 4 + ((size_t(chunk) * 2 + 15) & ~size_t(15)) + ((n_blocks * 12 + 15) & ~size_t(15));
}

template <bool SKIP_HETS, int WX>
__global__ void __launch_bounds__(GH_THREADS, 1) k_score_grouped_half(const GroupArgs a) {
    extern __shared__ __align__(16) unsigned char gh_smem[];
    const int wx = WX ? WX : a.wx;
    const int tw = 2 * wx;                            // threads per team
    const int spc = a.spc;
    const int q = threadIdx.x / tw, tt = threadIdx.x - q * tw;
    const int w = tt >> 1, h = tt & 1;
    const int slot = blockIdx.x * spc + q;
    const int word = blockIdx.y * wx + w;
    int seg = 0, begin = 0, end = 0;
    bool team_ok = false;
    if (q < spc) {
        const int j = a.jmax - 1 - slot / a.S, smp = slot % a.S;
        if (j >= 0 && j < a.seg_off[smp + 1] - a.seg_off[smp]) {
            team_ok = true;
            seg = a.seg_off[smp] + j;
            begin = a.mstart[smp] + j * a.chunk;
            end = min(a.mstart[smp + 1], begin + a.chunk);
        }
    }
    const int n_rows = end - begin;
    const int n_blocks = (n_rows + GH_BLOCK - 1) / GH_BLOCK;
    unsigned char *team = gh_smem + size_t(q < spc ? q : 0) * half_team_smem(wx, a.chunk);
    uint64_t *ring = reinterpret_cast<uint64_t *>(team);
    int32_t *s_row = reinterpret_cast<int32_t *>(team + size_t(GH_RING) * wx * 8);
    uint16_t *s_gid = reinterpret_cast<uint16_t *>(team + size_t(GH_RING) * wx * 8 + size_t(a.chunk) * 4);
    // s_chg[3 b + c]: bit k set <=> the weight of class c (ref, alt, het) at row 32 b + k differs from the row before it
    uint32_t *s_chg = reinterpret_cast<uint32_t *>(team + size_t(GH_RING) * wx * 8 + size_t(a.chunk) * 4 + ((size_t(a.chunk) * 2 + 15) & ~size_t(15)));
    if (team_ok) {
        for (int b = tt; b < 3 * n_blocks; b += tw) s_chg[b] = 0u;
#pragma unroll 4
        for (int r = tt; r < n_rows; r += tw) {
            s_row[r] = __ldg(a.pair_db + begin + r);
            s_gid[r] = __ldg(a.pair_gid + begin + r);
        }
    }
    __syncthreads();
    if (team_ok) {
        for (int r = tt + 1; r < n_rows; r += tw) {
            const int g1 = s_gid[r], g0 = s_gid[r - 1];
            if (g1 != g0) {
                const double4 t1 = reinterpret_cast<const double4 *>(a.table)[g1];
                const double4 t0 = reinterpret_cast<const double4 *>(a.table)[g0];
                const uint32_t bit = 1u << (r & 31);
                if (t1.x != t0.x) atomicOr(s_chg + 3 * (r >> 5), bit);
                if (t1.y != t0.y) atomicOr(s_chg + 3 * (r >> 5) + 1, bit);
                if (t1.z != t0.z) atomicOr(s_chg + 3 * (r >> 5) + 2, bit);
            }
        }
    }
    __syncthreads();                              // the last CTA-wide barrier
    if (!team_ok || word >= a.stride) return;

    const int64_t stride = a.stride;
    const int n_full = n_rows / GH_BLOCK;
    // four threads (the two halves of words 2i and 2i+1: always lanes of one warp) share the 16-byte pieces that hold both
    // words: thread `part` copies rows 8 part .. 8 part + 7 of the block
    const int part = ((w & 1) << 1) | h;
    const uint32_t pair_ring = smem_u32(ring + (w & ~1));
    const uint32_t ring_pitch = uint32_t(wx) * 8u;
    const unsigned char *pair_col = reinterpret_cast<const unsigned char *>(a.packed + (word & ~1));
    const uint32_t stride_b = uint32_t(stride) * 8u;
    auto issue = [&](int b) {
        const int r0 = b * GH_BLOCK + 8 * part;
        const uint32_t slot0 = pair_ring + uint32_t(r0 % GH_RING) * ring_pitch;
        if (b < n_full) {
#pragma unroll
            for (int k4 = 0; k4 < 8; k4 += 4) {
                const int4 rr = *reinterpret_cast<const int4 *>(s_row + r0 + k4);
                cp_async16(slot0 + uint32_t(k4 + 0) * ring_pitch, pair_col + (unsigned long long)(uint32_t(rr.x)) * stride_b);
                cp_async16(slot0 + uint32_t(k4 + 1) * ring_pitch, pair_col + (unsigned long long)(uint32_t(rr.y)) * stride_b);
                cp_async16(slot0 + uint32_t(k4 + 2) * ring_pitch, pair_col + (unsigned long long)(uint32_t(rr.z)) * stride_b);
                cp_async16(slot0 + uint32_t(k4 + 3) * ring_pitch, pair_col + (unsigned long long)(uint32_t(rr.w)) * stride_b);
            }
        } else if (b < n_blocks) {
            for (int k = 0; k < 8 && r0 + k < n_rows; ++k) cp_async16(slot0 + uint32_t(k) * ring_pitch, pair_col + (unsigned long long)(uint32_t(s_row[r0 + k])) * stride_b);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int b = 0; b < GH_INFLIGHT; ++b) issue(b);

    double F[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) F[b] = 0.0;
    BitCounter<GR_LP> c_int, c_ninfo, c_ref, c_alt, c_het;
    c_int.clear();
    c_ninfo.clear();
    c_ref.clear();
    c_alt.clear();
    c_het.clear();
    double w_ref, w_alt, w_het;
    {
        const double4 t = *reinterpret_cast<const double4 *>(a.table + 4 * size_t(s_gid[0]));
        w_ref = t.x;
        w_alt = t.y;
        w_het = t.z;
    }
    const uint32_t sel = h ? 0x7632u : 0x5410u;   // bytes of this thread's half of rows 2j (low) and 2j+1 (high)
    auto flush_class = [&](BitCounter<GR_LP> &c, double wt) {
        if (c.any()) {
            c_ninfo.add_counter(c);               // lanes are independent: the two halves add side by side
            if (wt == 1.0) c_int.add_counter(c);
            else if (wt != 0.0) {
                int32_t v[16];
                half_counts(c, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) F[i] = fma(wt, double(v[i]), F[i]);
            }
            c.clear();
        }
    };
    // mask: bit k = row k of the block starts a new weight of this class
    auto add_class = [&](BitCounter<GR_LP> &c, double &wt, const uint32_t (&pl)[16], uint32_t mask, int which, int r0) {
        if (mask == 0u) {
            c.add16(pl);
            return;
        }
        int k0 = 0;
        while (true) {
            const int k1 = mask ? __ffs(mask) - 1 : GH_BLOCK;
            if (k1 > k0) {
                const uint32_t rm = (k1 >= 32 ? 0xffffffffu : ((1u << k1) - 1u)) & ~((1u << k0) - 1u);      // rows k0 .. k1-1
                uint32_t m[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t even = uint32_t(int32_t(rm << (31 - 2 * j)) >> 31) & 0x0000FFFFu;
                    const uint32_t odd = uint32_t(int32_t(rm << (30 - 2 * j)) >> 31) & 0xFFFF0000u;
                    m[j] = pl[j] & (even | odd);
                }
                c.add16(m);
            }
            if (k1 >= GH_BLOCK) break;
            flush_class(c, wt);
            wt = a.table[4 * size_t(s_gid[r0 + k1]) + which];
            mask &= mask - 1u;
            k0 = k1;
        }
    };
    auto score_block = [&](const uint32_t (&lo)[16], const uint32_t (&hi)[16], int b) {
        const uint32_t m_ref = s_chg[3 * b], m_alt = s_chg[3 * b + 1], m_het = s_chg[3 * b + 2];
        uint32_t pl[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pl[j] = ~(lo[j] | hi[j]);
        add_class(c_ref, w_ref, pl, m_ref, 0, b * GH_BLOCK);
#pragma unroll
        for (int j = 0; j < 16; ++j) pl[j] = lo[j] & ~hi[j];
        add_class(c_alt, w_alt, pl, m_alt, 1, b * GH_BLOCK);
        if (!SKIP_HETS) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pl[j] = hi[j] & ~lo[j];
            add_class(c_het, w_het, pl, m_het, 2, b * GH_BLOCK);
        }
    };

    for (int b = 0; b < n_blocks; ++b) {
        if (b < n_full) cp_async_wait<GH_INFLIGHT - 1>();
        else cp_async_wait<0>();
        __syncwarp();
        const int r0 = b * GH_BLOCK;
        const uint64_t *slot = ring + size_t(r0 % GH_RING) * wx + w;
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint64_t v0 = ~0ull, v1 = ~0ull;      // rows past the end read as missing everywhere
            if (r0 + 2 * j < n_rows) v0 = slot[size_t(2 * j) * wx];
            if (r0 + 2 * j + 1 < n_rows) v1 = slot[size_t(2 * j + 1) * wx];
            lo[j] = prmt(uint32_t(v0), uint32_t(v1), sel);
            hi[j] = prmt(uint32_t(v0 >> 32), uint32_t(v1 >> 32), sel);
        }
        score_block(lo, hi, b);
        __syncwarp();
        if (b < n_full) issue(b + GH_INFLIGHT);
    }
    flush_class(c_ref, w_ref);
    flush_class(c_alt, w_alt);
    if (!SKIP_HETS) flush_class(c_het, w_het);

    int32_t vi[16], vn[16];
    half_counts(c_int, vi);
    half_counts(c_ninfo, vn);
    const int64_t o = int64_t(seg) * a.a_pad + word;
    const int64_t lane_pitch = a.stride;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int b = 16 * h + i;
        a.part_score[o + b * lane_pitch] = F[i];
        a.part_int[o + b * lane_pitch] = vi[i] | (vn[i] << 16);
    }
}

}  // namespace snpm
