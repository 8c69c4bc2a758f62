#!/usr/bin/env python
"""Timings of the BASELINE.json configurations that bench.py's single JSON line does not cover: `cross` on the full
1135 x 10.7 M panel (configs[2]), a dense sample (every panel row), and the 20 000-accession panel on one GPU
(configs[4] at N=1).  Writes one JSON object per line; run on a B200:  python scripts/measure_configs.py > out.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402

ge.build()
from snpmatch_b200 import lib, synth  # noqa: E402
from snpmatch_b200.core import genomes, snp_genotype, snpmatch  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        out.append(time.perf_counter() - t0)
    return float(np.median(out))


def cross_full():
    n_rows, n_acc = 10_700_000, 1135
    g = snp_genotype.Genotype.synthetic(n_rows, n_acc)
    pos, regions = synth.panel_positions(n_rows)
    s = synth.make_sample_fast(pos, regions, n_acc, 7, seed=777)
    gen = genomes.Genome("athaliana_tair10")
    cnt, off, n_w, _ = gen.window_layout(np.array(synth.TAIR10_CHRS), 300000)
    kmax = snpmatch.identity_kmax_table(4000, 0.02)
    b = lib.Batch(g.db, [0, len(s["pos"])], s["chr_ix"], s["pos"], s["wei"])
    res = {}

    def run():
        b.run_windows(False, 300000, cnt, off, n_w, kmax)
        b.epilogue()
        tot = b.fetch()
        res["w"] = b.fetch_window_rows()             # the surviving rows, compacted on the device
        top = np.argsort(-tot["prob"][0])[:10]
        res["f1"] = b.f1_pairs(top)
        res["tot"] = tot
    t = timed(run)

    def run_full():
        b.run_windows(False, 300000, cnt, off, n_w, kmax)
        b.epilogue()
        b.fetch()
        b.fetch_windows()                            # every window x accession cell (13 MB)
    t_full = timed(run_full)
    tm = b.timings()
    m = int(res["tot"]["m"][0])
    out = {"config": "configs[2]: cross, 399 windows of 300 kb + 45 simulated F1s, one PL sample (%d markers, %d matched) vs 1135 x 10.7M" % (len(s["pos"]), m),
           "host_call_ms": t * 1e3, "host_call_ms_full_window_arrays": t_full * 1e3, "surviving_rows": int(len(res["w"]["acc"])), "device_ms": tm["total_ms"], "score_kernel_ms": tm["score_ms"], "join_ms": tm["join_ms"],
           "comparisons_per_s_end_to_end": m * n_acc / t, "windows_with_markers": int((res["w"]["nrows"] > 0).sum()),
           "top_accession": int(np.nanargmin(res["tot"]["L"][0]))}
    b.close()
    # single-sample inbred latency through the host-buffer call
    def one():
        bb = g.db.scratch_batch([0, len(s["pos"])], s["chr_ix"], s["pos"], s["wei"])
        bb.run(); bb.epilogue(); bb.fetch()
    t1 = timed(one)
    out2 = {"config": "configs[1] single sample latency: one PL sample (%d matched) vs 1135 x 10.7M, host buffers in/out" % m,
            "host_call_ms": t1 * 1e3, "comparisons_per_s": m * n_acc / t1}
    # dense sample: every panel row
    rows = np.arange(n_rows)
    chrom = np.searchsorted(regions[:, 1], rows, side="right").astype(np.int32)
    rng = np.random.default_rng(5)
    wei = synth._pl_weights(rng, rng.integers(0, 2, size=n_rows).astype(np.int8), 1 + rng.poisson(3, size=n_rows))[1]
    bd = lib.Batch(g.db, [0, n_rows], chrom, pos, wei)
    outs = []
    for algo, name in ((lib.JOIN_SEARCH, "binary search"), (lib.JOIN_MERGEPATH, "merge-path")):
        def run_d():
            bd.run(join_algo=algo); bd.epilogue(); bd.wait()
        td = timed(run_d, reps=3, warm=1)
        tmd = bd.timings()
        outs.append({"config": "dense sample (m = N = 10.7M rows) vs 1135 accessions, join = %s" % name, "device_ms": tmd["total_ms"],
                     "join_ms": tmd["join_ms"], "score_kernel_ms": tmd["score_ms"], "combine_ms": tmd["combine_ms"],
                     "comparisons_per_s": n_rows * n_acc / td,
                     "score_GBps": (n_rows * (284 + 24) + 16 * n_acc) / (tmd["score_ms"] * 1e-3) / 1e9})
    bd.close()
    g.close()
    return [out, out2] + outs


def p20k():
    n_rows, n_acc = 10_700_000, 20000
    g = snp_genotype.Genotype.synthetic(n_rows, n_acc)
    pos, regions = synth.panel_positions(n_rows)
    S = 16
    samples = [synth.make_sample_fast(pos, regions, n_acc, (7 + 13 * i) % n_acc, seed=9000 + i) for i in range(S)]
    offs = np.concatenate([[0], np.cumsum([len(s["pos"]) for s in samples])])
    b = lib.Batch(g.db, offs, np.concatenate([s["chr_ix"] for s in samples]), np.concatenate([s["pos"] for s in samples]),
                  np.concatenate([s["wei"] for s in samples]))
    def run():
        b.run(); b.epilogue(); b.wait()
    t = timed(run, reps=3, warm=1)
    tm = b.timings()
    r = b.fetch()
    m = int(r["m"].sum())
    ok = all(int(np.nanargmin(r["L"][i])) == (7 + 13 * i) % n_acc for i in range(S))
    out = {"config": "configs[4] at N=1: %d PL samples vs the 20 000-accession x 10.7M panel (%.1f GB packed) on one B200" % (S, g.db.packed_bytes / 1e9),
           "device_ms": tm["total_ms"], "score_kernel_ms": tm["score_ms"], "comparisons_per_s": m * n_acc / t,
           "score_GBps": (m * (n_acc // 4 + 24) + 16 * n_acc * S) / (tm["score_ms"] * 1e-3) / 1e9, "true_accessions_recovered": ok}
    b.close()
    g.close()
    return [out]


if __name__ == "__main__":
    for rec in cross_full() + p20k():
        print(json.dumps(rec), flush=True)
