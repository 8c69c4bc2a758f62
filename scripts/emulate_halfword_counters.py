#!/usr/bin/env python
"""Bit-level NumPy emulation of the next design of k_score_grouped: a thread owns HALF a word column (16
accessions) and packs TWO rows per 32-bit register, so that the carry-save adders of the present kernel (BitCounter::add16,
csrc/grouped.cuh) count 32 rows per call on 32 useful bits, and the two halves of a counter are added at the read-out.
Checks, against direct counting, the three pieces whose bit layouts are easy to get wrong:
  1. the row-pair packing (PRMT selectors 0x5410 / 0x7632 on the low / high plane words of rows 2k and 2k+1),
  2. add16 on packed planes + the ripple into a 10-plane counter, over several blocks,
  3. the read-out: transpose_planes8 (byte j of t[i] <-> lane 8j+i), count(accession a) = lane a + lane a+16 via 16-bit fields,
     plus planes 8 and 9.
Runs on the CPU:  python scripts/emulate_halfword_counters.py"""
import numpy as np

U = np.uint32
M32 = U(0xFFFFFFFF)


def prmt(a, b, sel):
    """PTX prmt.b32 (default mode): result byte i = byte sel_i of the 8 bytes {a: 0-3, b: 4-7}."""
    src = [(int(a) >> (8 * k)) & 0xFF for k in range(4)] + [(int(b) >> (8 * k)) & 0xFF for k in range(4)]
    return U(sum(src[(sel >> (4 * i)) & 0x7] << (8 * i) for i in range(4)))


def csa(a, b, c):
    u = a ^ b
    return (a & b) | (u & c), u ^ c          # (carry, sum)


def add16(p, m):
    """BitCounter<P>::add16 of csrc/grouped.cuh: p = list of planes (uint32), m = 16 one-bit planes."""
    c0, s0 = csa(m[0], m[1], m[2]); c1, s1 = csa(m[3], m[4], m[5]); c2, s2 = csa(m[6], m[7], m[8])
    c3, s3 = csa(m[9], m[10], m[11]); c4, s4 = csa(m[12], m[13], m[14])
    c5, t0 = csa(s0, s1, s2); c6, t1 = csa(s3, s4, m[15]); c7, p[0] = csa(t0, t1, p[0])
    d0, u0 = csa(c0, c1, c2); d1, u1 = csa(c3, c4, c5); d2, u2 = csa(c6, c7, p[1]); d3, p[1] = csa(u0, u1, u2)
    e0, v0 = csa(d0, d1, d2); e1, p[2] = csa(v0, d3, p[2]); s, p[3] = csa(e0, e1, p[3])
    for k in range(4, len(p)):
        c = p[k] & s
        p[k] ^= s
        s = c
    assert s == 0, "counter overflow"


def transpose_planes8(r):
    def swap(a, b, s, m):
        t = ((r[a] >> U(s)) ^ r[b]) & U(m)
        r[b] ^= t
        r[a] ^= (t << U(s)) & M32
    for a, b in ((0, 1), (2, 3), (4, 5), (6, 7)):
        swap(a, b, 1, 0x55555555)
    for a, b in ((0, 2), (1, 3), (4, 6), (5, 7)):
        swap(a, b, 2, 0x33333333)
    for a, b in ((0, 4), (1, 5), (2, 6), (3, 7)):
        swap(a, b, 4, 0x0F0F0F0F)


def read_out(p):
    """counts of the 16 accessions of a half-word thread from a 10-plane counter whose lane a holds the even rows and lane
    a + 16 the odd rows of accession a."""
    t = [U(x) for x in p[:8]]
    transpose_planes8(t)
    out = np.zeros(16, dtype=np.int64)
    for i in range(8):
        x = t[i] & U(0x00FF00FF)                 # [lane i | lane 16 + i << 16]
        y = (t[i] >> U(8)) & U(0x00FF00FF)       # [lane 8 + i | lane 24 + i << 16]
        out[i] = int((x + (x >> U(16))) & U(0xFFFF))
        out[8 + i] = int((y + (y >> U(16))) & U(0xFFFF))
    for a in range(16):
        for k in (8, 9):
            out[a] += ((int(p[k]) >> a) & 1) << k
            out[a] += ((int(p[k]) >> (a + 16)) & 1) << k
    return out


def main():
    rng = np.random.default_rng(0)
    for trial in range(20):
        n_blocks = int(rng.integers(1, 12))                      # 32 rows per block; <= 352 rows: fits 10 planes per half
        n_rows = 32 * n_blocks
        code = rng.choice(4, size=(n_rows, 32), p=[0.55, 0.3, 0.05, 0.1])      # 0 ref, 1 alt, 2 het, 3 missing
        lo = np.array([sum(int(c & 1) << j for j, c in enumerate(row)) for row in code], dtype=np.uint32)
        hi = np.array([sum(int(c >> 1) << j for j, c in enumerate(row)) for row in code], dtype=np.uint32)
        for h in (0, 1):
            sel = 0x5410 if h == 0 else 0x7632
            counters = {k: [U(0)] * 10 for k in ("ref", "alt", "het")}
            for b in range(n_blocks):
                planes = {k: [] for k in counters}
                for j in range(16):
                    r0, r1 = 32 * b + 2 * j, 32 * b + 2 * j + 1
                    lp, hp = prmt(lo[r0], lo[r1], sel), prmt(hi[r0], hi[r1], sel)
                    assert int(lp) == ((int(lo[r0]) >> (16 * h)) & 0xFFFF) | (((int(lo[r1]) >> (16 * h)) & 0xFFFF) << 16)
                    planes["ref"].append(~(lp | hp) & M32)
                    planes["alt"].append(lp & ~hp & M32)
                    planes["het"].append(hp & ~lp & M32)
                for k in counters:
                    add16(counters[k], planes[k])
            for k, want_code in (("ref", 0), ("alt", 1), ("het", 2)):
                got = read_out(counters[k])
                want = (code[:, 16 * h:16 * h + 16] == want_code).sum(axis=0)
                assert np.array_equal(got, want), (trial, h, k, got, want)
    print("half-word / row-pair counters: packing, add16 on packed planes and the read-out agree with direct counts (20 trials)")


if __name__ == "__main__":
    main()
