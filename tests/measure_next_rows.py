#!/usr/bin/env python
"""Timings of the SURVEY 8(f) rows built next to the matching path, on the full-size synthetic panel (1135 x 10.7 M):
whole-column reads of the resident panel, `pairsnp`, and the `genotype_cross` window genotyper, each beside the CPU oracle on a
bounded sample.  Lives under tests/ because it runs the CPU oracle next to the GPU path (only tests/, smoke() and bench.py's CPU
legs may use oracle/); it is a measurement script, not collected by pytest.  One JSON object per line; run on a B200:
python tests/measure_next_rows.py > out.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402

ge.build()
from oracle import snpmatch_oracle as orc  # noqa: E402  (CPU baseline leg only)
from snpmatch_b200 import lib, synth  # noqa: E402
from snpmatch_b200.core import genomes, genotype_cross, snp_genotype  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        out.append(time.perf_counter() - t0)
    return float(np.median(out))


def main():
    n_rows, n_acc = 10_700_000, 1135
    g = snp_genotype.Genotype.synthetic(n_rows, n_acc)
    pos, regions = synth.panel_positions(n_rows)
    row_chr = np.searchsorted(regions[:, 0], np.arange(n_rows), side="right") - 1
    rng = np.random.default_rng(11)

    # ---- whole accession columns -------------------------------------------------------------------------
    cols = np.array([3, 21])
    t = timed(lambda: g.g.db.read_columns(cols))
    t16 = timed(lambda: g.g.db.read_columns(np.arange(16) * 64), reps=3, warm=1)
    sample_rows = 1_000_000
    host_panel = synth.panel_codes(synth.SEED_PANEL, np.arange(sample_rows), n_acc)
    t_cpu = timed(lambda: np.ascontiguousarray(host_panel[:, cols].T), reps=3, warm=1) * n_rows / sample_rows
    print(json.dumps({"config": "8(f): two accession columns of the resident 1135 x 10.7M panel (g_acc.snps[:, [i, j]]), host call incl. D2H",
                      "host_call_ms": t * 1e3, "sector_GBps": n_rows * 32 / t / 1e9, "d2h_bytes": int(2 * n_rows),
                      "sixteen_columns_in_sixteen_sectors_ms": t16 * 1e3,
                      "cpu_numpy_int8_ndarray_ms": t_cpu * 1e3, "cpu_sample": "1M-row int8 ndarray in RAM, scaled to 10.7M rows"}), flush=True)

    # ---- genotype_cross ----------------------------------------------------------------------------------------
    p = g.g.db.read_columns(cols)
    seg = orc.segregating_parent_markers(p[0], p[1])
    par_chr = np.array(synth.TAIR10_CHRS)[row_chr[seg]]
    par_pos, p1, p2 = pos[seg].astype(np.int64), p[0][seg], p[1][seg]
    S, n_vcf = 384, 200_000
    rows = np.sort(rng.choice(n_rows, n_vcf, replace=False))
    vcf_chr, vcf_pos = np.array(synth.TAIR10_CHRS)[row_chr[rows]], pos[rows].astype(np.int64)
    codes = rng.choice(np.array([-1, 0, 1, 2], dtype=np.int8), size=(n_vcf, S), p=[0.3, 0.35, 0.2, 0.15])
    gen = genomes.Genome("athaliana_tair10")
    res = {}

    def run():
        res["r"] = genotype_cross.window_calls(par_chr, par_pos, p1, p2, vcf_chr, vcf_pos, codes, gen, 300000, 1.5)
    t = timed(run, reps=3, warm=1)
    r = res["r"]
    m = int(r["n_matched"].sum())
    # the kernel call alone (inputs prepared): pairs ordered by window
    from snpmatch_b200.core.snp_genotype import Genotype
    i_p, i_v = Genotype.get_common_positions(par_chr, par_pos, vcf_chr, vcf_pos)
    win = np.repeat(np.arange(len(r["n_matched"])), r["n_matched"])
    ws = np.concatenate([[0], np.cumsum(r["n_matched"])]).astype(np.int32)
    t_k = timed(lambda: lib.cross_window_genotypes(i_p, i_v, ws, p1, p2, codes, 1.5), reps=3, warm=1)
    # CPU oracle on a bounded sample: 8 samples
    names = np.array(["./.", "0/0", "1/1", "0/1"])
    sub = names[codes[:, :8].astype(int) + 1]
    t0 = time.perf_counter()
    calls, counts, n_matched = orc.genotype_cross_windows(par_chr, par_pos, p1, p2, vcf_chr, vcf_pos, sub, gen.chrs, gen.chrlen, 300000, 1.5)
    t_cpu = time.perf_counter() - t0
    ok = bool(np.array_equal(r["n_matched"], n_matched) and all(np.array_equal(r["counts"][w][:8], c) for w, c in counts.items())
              and all([(-1 if x == "NA" else x) for x in calls[w]] == r["geno"][w][:8].tolist() for w in counts))
    print(json.dumps({"config": "8(f): genotype_cross, %d samples x %d VCF markers, parents 3x21 of the 1135 x 10.7M panel (%d segregating markers, %d matched), %d windows of 300 kb"
                                % (S, n_vcf, len(seg), m, len(r["n_matched"])),
                      "host_call_ms": t * 1e3, "window_kernel_call_ms_incl_h2d": t_k * 1e3, "h2d_bytes": int(codes.nbytes + 16 * m + 2 * len(seg)),
                      "marker_sample_comparisons_per_s": m * S / t_k, "window_sample_cells_per_s": len(r["n_matched"]) * S / t_k,
                      "cpu_oracle_s_for_8_samples": t_cpu, "cpu_oracle_window_sample_cells_per_s": len(r["n_matched"]) * 8 / t_cpu,
                      "parity_first_8_samples": ok, "borderline_cells": int(r["borderline"].sum())}), flush=True)

    # ---- pairsnp ------------------------------------------------------------------------------------------------
    n1 = n2 = 1_000_000
    r1, r2 = np.sort(rng.choice(n_rows, n1, replace=False)), np.sort(rng.choice(n_rows, n2, replace=False))
    c1, q1 = np.array(synth.TAIR10_CHRS)[row_chr[r1]], pos[r1].astype(np.int64)
    c2, q2 = np.array(synth.TAIR10_CHRS)[row_chr[r2]], pos[r2].astype(np.int64)
    gts = np.array(["0/0", "1/1", "0/1"])
    g1, g2 = gts[rng.integers(0, 3, n1)], gts[rng.integers(0, 3, n2)]
    d = os.path.join("/tmp", "snpm_pairsnp_%d" % os.getpid())
    os.makedirs(d, exist_ok=True)
    for name, c, q, gt in (("a", c1, q1, g1), ("b", c2, q2, g2)):
        np.savez(os.path.join(d, name + ".npz"), chr=c, pos=q, gt=gt, wei=np.ones((len(q), 3)), dp=np.ones(len(q)))
    from snpmatch_b200.core import snpmatch
    st = {}

    def run_pair():
        st["r"] = snpmatch.pairwiseScore(os.path.join(d, "a.npz"), os.path.join(d, "b.npz"), False, outFile=None, hdf5File=g)
    t = timed(run_pair, reps=3, warm=1)
    t0 = time.perf_counter()
    want = orc.pairwise_score(c1, q1, g1, c2, q2, g2, "a.npz", "b.npz")
    t_cpu = time.perf_counter() - t0
    got = st["r"]
    ok = all(got[k][1] == want[k][1] and got[k][0] == want[k][0] for k in ("1", "2", "3", "4", "5", "matches"))
    print(json.dumps({"config": "8(f): pairsnp, two samples of 1M markers on panel positions, restricted to the resident panel (two joins + counting)",
                      "host_call_ms_incl_npz_load": t * 1e3, "common_markers": int(got["matches"][1]), "cpu_oracle_ms_without_panel_join": t_cpu * 1e3,
                      "parity": bool(ok)}), flush=True)
    g.close()


if __name__ == "__main__":
    main()
