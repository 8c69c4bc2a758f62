"""Development measurement of the coded path (device-side grouping + k_score_grouped2) on the headline workload
(64 PL samples x 50 k markers vs 1135 x 10.7 M): device times per stage (library CUDA events) and parity against the
order-exact kernel; the host-grouped kernel of round 1 next to it.
    python scripts/measure_coded.py [--samples 64] [--chunks 320,352] [--accessions 1135] [--shard-of K]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--rows", type=int, default=10_700_000)
    ap.add_argument("--accessions", type=int, default=1135)
    ap.add_argument("--markers", type=int, default=45000)
    ap.add_argument("--chunks", default="320")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--no-exact", action="store_true")
    ap.add_argument("--no-old", action="store_true")
    ap.add_argument("--hard", action="store_true")
    ap.add_argument("--shard-of", type=int, default=1)
    ap.add_argument("--lib", default="", help="another build of the library (A/B runs in one session)")
    args = ap.parse_args()
    if args.lib:
        os.environ["SNPM_LIB_PATH"] = os.path.abspath(args.lib)
    import __graft_entry__ as ge
    ge.build()
    from snpmatch_b200 import lib, synth
    out_lib = lib.LIB_PATH
    from snpmatch_b200.core import snp_genotype
    import bench
    positions, regions = synth.panel_positions(args.rows)
    g = snp_genotype.Genotype.synthetic(args.rows, args.accessions, device=0, row_range=(0, args.rows // args.shard_of))
    db = g.db
    bench.N_EXTRA_MARKERS = 5000
    samples = bench.make_samples(positions, regions, args.accessions, args.samples, args.markers)
    offs = np.concatenate([[0], np.cumsum([len(s["pos"]) for s in samples])]).astype(np.int64)
    chrom = np.concatenate([s["chr_ix"] for s in samples]).astype(np.int32)
    pos = np.concatenate([s["pos"] for s in samples]).astype(np.int32)
    wei = np.concatenate([synth.hard_weights(s["code"]) if args.hard else s["wei"] for s in samples])
    b = lib.Batch(db, offs, chrom, pos, wei)
    out = {"samples": args.samples, "accessions": args.accessions, "shard_of": args.shard_of, "lib": out_lib}
    exact = None
    if not args.no_exact:
        b.run(kernel_mode=lib.KERNEL_POPCOUNT if args.hard else lib.KERNEL_FP64)
        b.epilogue()
        b.wait()
        exact = {k: v.copy() for k, v in b.fetch().items()}
        out["exact"] = b.timings()
    t0 = time.perf_counter()
    cs = lib.code_markers(offs, chrom, pos, wei)
    out["code_markers_host_s"] = time.perf_counter() - t0
    out["distinct_weights"] = int(len(cs.wtable))

    def summarise(ts, r, guard):
        t = {k: float(np.median([x[k] for x in ts])) for k in ts[0]}
        m_total = int(r["m"].sum())
        alg = m_total * ((args.accessions + 3) // 4 + 24) + 16 * args.accessions * args.samples
        t["score_GBps"] = alg / (t["score_ms"] * 1e-3) / 1e9
        t["frac_of_6526"] = t["score_GBps"] / 6526.5
        t["guard_flagged_samples"] = int((guard > 0).sum())
        if exact is not None:
            ok = guard == 0
            t["matches_equal"] = bool(np.array_equal(r["matches"][ok], exact["matches"][ok]))
            t["ninfo_equal"] = bool(np.array_equal(r["ninfo"], exact["ninfo"]))
            t["score_max_rel"] = float(np.max(np.abs(r["score"] - exact["score"]) / np.maximum(exact["score"], 1.0)))
        return t

    for chunk in [int(c) for c in args.chunks.split(",")]:
        b.set_group_chunk(chunk)
        b.upload_coded(cs)
        ts = []
        for _ in range(args.reps):
            b.run(kernel_mode=lib.KERNEL_GROUPED)
            b.epilogue()
            b.wait()
            t = b.coded_timings()
            t["total_ms"] = b.timings()["total_ms"]
            ts.append(t)
        out["coded_chunk_%d" % chunk] = summarise(ts, b.fetch(), b.guard_counts())
    if not args.no_old:
        t0 = time.perf_counter()
        gs = lib.group_markers(offs, chrom, pos, wei)
        out["group_markers_host_s"] = time.perf_counter() - t0
        b.set_group_chunk(320)
        b.upload_grouped(gs)
        ts = []
        for _ in range(args.reps):
            b.run(kernel_mode=lib.KERNEL_GROUPED)
            b.epilogue()
            b.wait()
            ts.append(b.timings())
        out["host_grouped_chunk_320"] = summarise(ts, b.fetch(), b.guard_counts())
    print(json.dumps(out, indent=1))
    b.close()
    g.close()


if __name__ == "__main__":
    main()
