#!/usr/bin/env python
"""
bench.py — SNP x accession comparisons/s of the genotype-matching hot path (BASELINE.json metric).

Workload (config.workload): BASELINE configs[1] — the synthetic 1001-Genomes-shaped panel
(1135 accessions x 10.7 M SNPs, generated in HBM) scored against low-coverage samples with PL
weights (~50 k markers each, ~45 k of them in the panel).  One step = one pass of the hot path
(join + chunked scoring + combine + likelihood epilogue) over one batch of `--samples` independent
samples, each scored on its own exactly as separate `snpmatch inbred` runs would.

  value      comparisons/s with the batch already resident in HBM (device-timed, CUDA events on the
             stream the kernels run on, max over ranks);
  e2e        the same through the host-buffer API: per step the H2D copy of the samples from pinned
             memory and the D2H read of scores / counts / likelihoods are inside the timed region;
  roofline   the scoring kernel (k_score_grouped, the grouped counting kernel): algorithmic bytes per launch / its
             CUDA-event time; the order-exact fp64 kernel (k_score_segments) is reported next to it;
  cpu_baseline  the CPU oracle (a NumPy restatement of the reference path) on one sample, one core.

N > 1 (torchrun): the panel is sharded by SNP-row ranges, samples are replicated, per-GPU partial
scores/counts are summed with one reduce-scatter (a one-shot pull over peer memory, or NCCL with --reduce nccl), then every rank runs the epilogue on, and reads back, its
share of the samples.
The batch grows with N (samples = N x --samples) so that per-GPU work is fixed: "scaling": "weak".

`--impl reference` times the reference's own CPU path (oracle port; one process per sample on all
host cores, the way the reference is deployed) for the same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ROWS = 10_700_000
N_ACC = 1135
N_DB_MARKERS = 45_000
N_EXTRA_MARKERS = 5_000
E2E_DEPTH = 3            # batch objects rotating in the end-to-end arm
METRIC = "snp_accession_comparisons_per_sec"
UNIT = "comparisons/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=64, help="samples per step and per GPU")
    ap.add_argument("--rows", type=int, default=N_ROWS)
    ap.add_argument("--accessions", type=int, default=N_ACC)
    ap.add_argument("--markers", type=int, default=N_DB_MARKERS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--group-chunk", type=int, default=320, help="rows per segment of the grouped kernel")
    ap.add_argument("--force-exact", action="store_true", help="order-exact fp64 kernel as the headline path")
    ap.add_argument("--reduce", default="auto", choices=["auto", "p2p", "nccl", "none"],
                    help="cross-GPU sum of the per-sample totals: p2p = one-shot reduce over peer memory (CUDA IPC over NVLink, flag barrier + "
                         "pulls in one kernel), nccl = NCCL reduce-scatter, auto = what was measured faster (p2p on 2 GPUs; NCCL beyond, "
                         "where it reduces inside the NVSwitch), none = timing experiment only (totals stay per rank)")
    ap.add_argument("--cpu-markers", type=int, default=0, help="bound the CPU sample (0 = one whole sample)")
    args = ap.parse_args()
    if args.reduce == "auto":
        # measured: 2 GPUs 0.501 ms/step p2p vs 0.505 NCCL; 8 GPUs 0.745 p2p vs 0.720 NCCL
        args.reduce = "p2p" if int(os.environ.get("WORLD_SIZE", args.gpus)) == 2 else "nccl"
    return args


def workload_config(args, n_gpus, n_samples):
    return {
        "workload": "configs[1]: inbred scoring of %d low-coverage PL samples (~%d markers each, %d in the panel) against the "
                    "synthetic 1001G-shaped panel %d accessions x %d SNPs" % (
                        n_samples, args.markers + N_EXTRA_MARKERS, args.markers, args.accessions, args.rows),
        "panel_rows": args.rows, "accessions": args.accessions, "samples_per_step": n_samples,
        "markers_per_sample": args.markers + N_EXTRA_MARKERS, "weights": "PL (exp(-PL/10), f64)", "kernel": "grouped counting kernel (markers ordered by weight triple at parse time)",
        "sharding": "single GPU" if n_gpus == 1 else "SNP-row ranges over %d GPUs + one reduce-scatter of the per-accession partials per step (%s; every rank finishes and reads back its share of the samples)" % (
            n_gpus, "one kernel over peer memory: flag barrier + 16-byte pulls through NVLink, no collective library on the path" if args.reduce == "p2p" else "NCCL"),
        "cache": "inputs larger than L2: each step gathers %.0f MB of distinct panel rows" % (
            n_samples * args.markers * ((args.accessions + 63) // 64 * 16) / 1e6),
    }


def make_samples(positions, regions, n_acc, n_samples, n_markers, first_seed=5000):
    from snpmatch_b200 import synth
    out = []
    for i in range(n_samples):
        out.append(synth.make_sample_fast(positions, regions, n_acc, true_acc=(7 + 13 * i) % n_acc, n_db=n_markers,
                                          n_extra=N_EXTRA_MARKERS, seed=first_seed + i))
    return out


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, one sample
# ------------------------------------------------------------------------------------------------------
def cpu_prepare_sample(sample, n_acc, max_markers=0):
    """Materialise (outside the timed region) the panel rows the sample touches, as the in-RAM int8 matrix the
    reference would read from HDF5."""
    from snpmatch_b200 import synth
    rows = sample["rows"][sample["rows"] >= 0]
    if max_markers and len(rows) > max_markers:
        rows = rows[:max_markers]
    codes = synth.panel_codes(synth.SEED_PANEL, rows, n_acc)
    return rows, codes


def cpu_run_sample(positions, regions, sample, rows, codes):
    """Timed region of the CPU arm: Genotyper.genotyper of the reference (snpmatch.py:207-233) as restated by
    the oracle — the (chrom,pos) join over the database's per-row chromosome labels, then 1000-row chunks of
    matchGTsAccs — plus the likelihood epilogue.  Returns (comparisons, seconds)."""
    from oracle import snpmatch_oracle as orc
    from snpmatch_b200 import synth
    n_acc = codes.shape[1]
    names = np.array(synth.TAIR10_CHRS)
    s_chrs = np.char.add("Chr", names[sample["chr_ix"]])
    s_pos = sample["pos"].astype(np.int64)
    lookup = np.full(len(positions), -1, dtype=np.int64)
    lookup[rows] = np.arange(len(rows))
    t0 = time.perf_counter()
    labels = orc.db_chromosome_labels(names, regions)                 # pygwas/genotype.py:156-161
    common = orc.get_common_positions(labels, positions, s_chrs, s_pos)
    keep = lookup[common[0]] >= 0                                     # bounded sample: only the materialised rows
    c0, c1 = common[0][keep], common[1][keep]
    score = np.zeros(n_acc)
    ninfo = np.zeros(n_acc, dtype=np.int64)
    for j in range(0, len(c0), 1000):
        t_s, t_n = orc.match_gts_accs(sample["wei"][c1[j:j + 1000]], codes[lookup[c0[j:j + 1000]]])
        score = score + t_s
        ninfo = ninfo + t_n
    with np.errstate(all="ignore"):
        orc.calculate_likelihoods(score.astype(np.int64), ninfo)
    dt = time.perf_counter() - t0
    return len(c0) * n_acc, dt, score, ninfo


_W = {}


def _worker(i):
    s = _W["samples"][i]
    comps, dt, _, _ = cpu_run_sample(_W["positions"], _W["regions"], s, *_W["prepared"][i])
    return comps, dt


def _prepare(i):
    return cpu_prepare_sample(_W["samples"][i], _W["n_acc"], _W["cap"])


def run_reference_arm(args):
    """--impl reference: one process per sample on all host cores (README.md:9 deployment model)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from snpmatch_b200 import synth
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, 32))
    positions, regions = synth.panel_positions(args.rows)
    samples = make_samples(positions, regions, args.accessions, procs, args.markers)
    cap = args.cpu_markers or 0
    ctx = mp.get_context("fork")
    _W.update(positions=positions, regions=regions, samples=samples, n_acc=args.accessions, cap=cap)
    with ctx.Pool(procs) as pool:                       # untimed: materialise the panel rows every sample touches
        prepared = pool.map(_prepare, range(procs))
    _W.update(prepared=prepared)                        # the timed pool below is forked after this, so workers inherit it
    times = []
    comps = 0
    with ctx.Pool(procs) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_worker, range(procs))
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
                comps = sum(r[0] for r in res)
    total = sum(times)
    value = comps * len(times) / total if total > 0 else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(len(times), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus, args.samples * args.gpus),      # the b200 arm's workload; a step here is a bounded sample of it
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": "%d samples of that batch per step, one process each (join over the %d database labels + %d-marker "
                                   "chunked matchGTsAccs + likelihoods); NumPy oracle port of snpmatch.py:207-233, database "
                                   "rows held in RAM as int8" % (procs, args.rows, cap or args.markers)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_b200_arm(args):
    import torch
    import __graft_entry__ as ge
    ge.build()
    from snpmatch_b200 import lib, sharding, synth
    from snpmatch_b200.core import snp_genotype

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert lib.device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_pg = dist.new_group(backend="gloo")      # host-side agreement on the (rare) samples to re-score: must not queue behind kernels
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)

    n_rows, n_acc = args.rows, args.accessions
    S = args.samples * world
    positions, regions = synth.panel_positions(n_rows)
    r0, r1 = sharding.shard_rows(n_rows, world, rank)
    g = snp_genotype.Genotype.synthetic(n_rows, n_acc, row_range=(r0, r1), device=local_rank)
    db = g.db
    db.set_stream(stream.cuda_stream)
    samples = make_samples(positions, regions, n_acc, S, args.markers)
    # a rank only needs the markers that can fall into its row range (constant per-GPU join work and H2D bytes)
    parts = samples
    slices = [(0, len(s["pos"])) for s in samples]
    if world > 1:
        parts, slices = [], []
        for s in samples:
            i0, i1 = sharding.shard_marker_range(s["chr_ix"], s["pos"], regions, positions, r0, r1)
            parts.append({k: s[k][i0:i1] for k in ("chr_ix", "pos", "wei")})
            slices.append((i0, i1))
    offs = np.concatenate([[0], np.cumsum([len(s["pos"]) for s in parts])]).astype(np.int64)
    n_tot = int(offs[-1])

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    arrs = []
    for a in (offs, np.concatenate([s["chr_ix"] for s in parts]).astype(np.int32),
              np.concatenate([s["pos"] for s in parts]).astype(np.int32),
              np.concatenate([s["wei"] for s in parts]).astype(np.float64)):
        t, v = pinned(a)
        keep.append(t)
        arrs.append(v)
    h_off, h_chr, h_pos, h_wei = arrs
    S_loc = S // world                                 # samples whose results this rank finishes and reads back
    out_t = {k: torch.empty((S_loc, n_acc), dtype=torch.float64 if k in ("score", "prob", "L", "LR") else torch.int64).pin_memory()
             for k in ("score", "matches", "ninfo", "prob", "L", "LR")}
    out_t["m"] = torch.empty(S_loc, dtype=torch.int64).pin_memory()
    out = {k: v.numpy() for k, v in out_t.items()}

    batch = lib.Batch(db, h_off, h_chr, h_pos, h_wei)       # position order: the order-exact fp64 kernel
    # grouped order (done once, at parse time): markers of every sample ordered by weight triple -> counting kernel
    t_group = time.perf_counter()
    gs_raw = lib.group_markers(h_off, h_chr, h_pos, h_wei)
    t_group = time.perf_counter() - t_group
    assert gs_raw is not None, "the synthetic PL weights qualify for the grouped kernel"
    g_arrs = []
    for a in (gs_raw.chrom, gs_raw.pos, gs_raw.gid, gs_raw.table, gs_raw.packed, gs_raw.run_gid, gs_raw.run_end):
        if a is None:
            g_arrs.append(None)
            continue
        t, v = pinned(a)
        keep.append(t)
        g_arrs.append(v)
    gs = lib.GroupedSamples(h_off, g_arrs[0], g_arrs[1], g_arrs[2], g_arrs[3], gs_raw.order, packed=g_arrs[4], run_gid=g_arrs[5], run_end=g_arrs[6])
    if os.environ.get("SNPM_BENCH_NO_RUNS"):           # experiment: ids as one uint16 per marker (6 bytes per marker) instead of runs
        gs.run_gid = gs.run_end = None
    gbatch = lib.Batch(db, h_off, h_chr, h_pos, h_wei)
    gbatch.set_group_chunk(args.group_chunk)
    gbatch.upload_grouped(gs)

    # Row sharding cuts every weight group into `world` pieces, which makes the counter read-outs more frequent; measured at
    # 4 and 8 GPUs the grouped kernel still wins (1.6e13 vs 1.2e13, 2.9e13 vs 2.3e13), so it is the headline path everywhere.
    rows_per_group = args.markers / float(world) / max(1, len(np.unique(samples[0]["wei"], axis=0)))
    use_grouped = not args.force_exact

    def reduce_totals(b):
        # reduce-scatter of the per-sample totals: every rank is left with the totals of its S/world samples and finishes
        # (epilogue) and reads back only those.  p2p: one kernel that is barrier + pull over peer memory (k_reduce_peers)
        if args.reduce == "p2p":
            sharding.p2p_reduce_scatter(b)
        elif args.reduce == "none":
            pass
        else:
            sharding.reduce_scatter_batch(b, dist, dev, rank, world)

    def run_batch(b, **kw):
        if world > 1 and args.reduce == "p2p":
            sharding.p2p_before_run(b, dist, rank, world, host_pg)      # maps the peers' buffers on first use (host collective)
        b.run(**kw)

    def own_share(b):
        if world > 1:
            b.set_result_range(rank * S_loc, S_loc)
        return b

    own_share(batch)
    own_share(gbatch)

    def device_step(b=None, mode=None):
        b = (gbatch if use_grouped else batch) if b is None else b
        run_batch(b, kernel_mode=lib.KERNEL_GROUPED if b is gbatch else lib.KERNEL_FP64)
        if world > 1:
            reduce_totals(b)
        b.epilogue()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    score_ms, total_launches = [], 0
    with torch.cuda.stream(stream):
        # ---- resident arm -------------------------------------------------------------------------
        head = gbatch if use_grouped else batch
        for _ in range(args.warmup):
            device_step()
        head.wait()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()                        # before the barrier: forking nvidia-smi takes ~1 ms, which the other ranks would
        barrier()                                  # otherwise spend waiting for rank 0 inside the first timed step
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            device_step()
        ev1.record(stream)
        head.wait()
        barrier()
        dev_ms = ev0.elapsed_time(ev1)
        # per-kernel times of the scoring kernel: one more timed pass that reads the library's own events each step
        stage_ms = {}
        for _ in range(args.steps):
            device_step()
            t = head.timings()
            score_ms.append(t["score_ms"])
            for k, v in t.items():
                stage_ms.setdefault(k, []).append(v)
            total_launches = t["launches"]
        guard_resident = head.guard_counts()
        barrier()
        # ---- the order-exact fp64 kernel on the same samples (position order), for comparison
        exact_kernel_ms = []
        for _ in range(args.warmup):
            device_step(batch)
        batch.wait()
        barrier()
        xv0, xv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xv0.record(stream)
        for _ in range(args.steps):
            device_step(batch)
        xv1.record(stream)
        batch.wait()
        barrier()
        exact_ms = xv0.elapsed_time(xv1)
        for _ in range(args.steps):
            device_step(batch)
            exact_kernel_ms.append(batch.timings()["score_ms"])
        exact_res = {k: v.copy() for k, v in batch.fetch().items()}
        barrier()
        # ---- called-genotype variant of the same samples (0/1 weights, as BED / GT-only VCF inputs give): popcount kernel
        hard = lib.Batch(db, h_off, h_chr, h_pos, np.concatenate([synth.hard_weights(s["code"][:len(p["pos"])] if world == 1 else
                                                                                     s["code"][sl[0]:sl[1]])
                                                                  for s, p, sl in zip(samples, parts, slices)]))
        own_share(hard)
        hard_ms, hard_kernel_ms = 0.0, []

        def hard_step():
            run_batch(hard, kernel_mode=lib.KERNEL_POPCOUNT)
            if world > 1:
                reduce_totals(hard)
            hard.epilogue()
        for _ in range(args.warmup):
            hard_step()
        hard.wait()
        barrier()
        hv0, hv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hv0.record(stream)
        for _ in range(args.steps):
            hard_step()
        hv1.record(stream)
        hard.wait()
        barrier()
        hard_ms = hv0.elapsed_time(hv1)
        for _ in range(args.steps):
            hard_step()
            hard_kernel_ms.append(hard.timings()["score_ms"])
        hard_res = hard.fetch()
        # ... and through the grouped counting kernel (three weight groups per sample: pure counting, no fp64 at all)
        hard_wei = np.concatenate([synth.hard_weights(s["code"][:len(p["pos"])] if world == 1 else s["code"][sl[0]:sl[1]])
                                   for s, p, sl in zip(samples, parts, slices)])
        hard_gs = lib.group_markers(h_off, h_chr, h_pos, hard_wei)
        hard.set_group_chunk(args.group_chunk)
        hard.upload_grouped(hard_gs)

        def hardg_step():
            run_batch(hard, kernel_mode=lib.KERNEL_GROUPED)
            if world > 1:
                reduce_totals(hard)
            hard.epilogue()
        for _ in range(args.warmup):
            hardg_step()
        hard.wait()
        barrier()
        gv0, gv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gv0.record(stream)
        for _ in range(args.steps):
            hardg_step()
        gv1.record(stream)
        hard.wait()
        barrier()
        hardg_ms = gv0.elapsed_time(gv1)
        hardg_kernel_ms = []
        for _ in range(args.steps):
            hardg_step()
            hardg_kernel_ms.append(hard.timings()["score_ms"])
        hardg_res = hard.fetch()
        hard_same = all(np.array_equal(hardg_res[k], hard_res[k], equal_nan=True) for k in ("score", "matches", "ninfo", "m", "L", "LR"))
        hard.close()
        barrier()
        # ---- end-to-end arm: host buffers in, host buffers out -------------------------------------
        # Every step uploads its inputs from pinned host memory (on the batch's own copy stream) and reads its results back.
        # E2E_DEPTH batch objects rotate.  One batch's cycle is upload (0.36 ms at 53 GB/s) -> kernels (0.43) -> read-back (0.07) ->
        # the host sees the results and uploads again; with two batches that cycle (1.1 ms with host latencies) bounds the step at
        # 0.56 ms, with three the kernels do.
        extra = []
        for _ in range(E2E_DEPTH - 1):
            bx = lib.Batch(db, h_off, h_chr, h_pos, h_wei)
            bx.set_group_chunk(args.group_chunk)
            own_share(bx)
            extra.append(bx)
        pair = [gbatch] + extra
        rescored = [0]

        coded = None if use_grouped else lib.index_weights(h_wei)
        if coded is not None:
            idx_t, h_idx = pinned(coded[0])
            tab_t, h_tab = pinned(coded[1])
            keep.extend([idx_t, tab_t])

        skip_part = os.environ.get("SNPM_E2E_SKIP", "")    # timing experiments: leave the upload or the read-back out of the steady state
        primed = set()

        def up(bt):
            if skip_part == "upload" and id(bt) in primed:
                return
            primed.add(id(bt))
            if use_grouped:
                bt.upload_grouped(gs)                        # ~4.1 bytes per marker
            elif coded is not None:
                bt.upload_indexed(h_off, h_chr, h_pos, h_idx, h_tab)      # 14 bytes per marker
            else:
                bt.upload(h_off, h_chr, h_pos, h_wei)

        # software pipeline over the rotating batches: while step k's results travel to the host and the host looks at them, the
        # kernels of the next steps are already queued and the samples of step k+depth are being copied in.  Every step still
        # uploads its own inputs (pinned host arrays) and reads back its own results (pinned host arrays) inside the timed region.
        outs = [dict(out)]
        for _ in range(E2E_DEPTH - 1):
            ot = {k: torch.empty_like(v).pin_memory() for k, v in out_t.items()}
            keep.append(ot)
            outs.append({k: v.numpy() for k, v in ot.items()})
        for o in outs:
            gt_ = torch.zeros(S, dtype=torch.int32).pin_memory()
            keep.append(gt_)
            o["guard"] = gt_.numpy()

        def launch(k):
            b_ = pair[k % E2E_DEPTH]
            run_batch(b_, kernel_mode=lib.KERNEL_GROUPED if use_grouped else lib.KERNEL_FP64)
            if world > 1:
                reduce_totals(b_)
            b_.epilogue()
            if skip_part == "fetch":
                b_.fetch_async({"m": outs[k % E2E_DEPTH]["m"], "guard": outs[k % E2E_DEPTH]["guard"]})      # the counts only (a few hundred bytes)
            else:
                b_.fetch_async(outs[k % E2E_DEPTH])                  # D2H of step k, queued behind its kernels

        def finish(k):
            b_ = pair[k % E2E_DEPTH]
            r = b_.fetch_wait()                              # results of step k (this rank's share) are on the host
            flagged = np.flatnonzero(r["guard"])             # int(score) needs the reference's summation order (~1e-4 per sample)
            if world > 1:
                # re-scoring a sample is a collective job (every rank holds a row range of the panel): the ranks agree on the
                # set once, at the end of the run (resolve_flagged), instead of paying a host collective every step
                pending.extend((k, int(rank * S_loc + sidx)) for sidx in flagged)
            else:
                for sidx in flagged:
                    rescore(int(sidx), r)
            return r

        pending = []

        def rescore(sidx, r=None):
            lo, hi = int(h_off[sidx]), int(h_off[sidx + 1])
            one = db.scratch_batch([0, hi - lo], h_chr[lo:hi], h_pos[lo:hi], h_wei[lo:hi])
            one.run()
            if world > 1:
                sharding.allreduce_batch(one, dist, dev)
            one.epilogue()
            r1 = one.fetch()
            if r is not None and sidx // S_loc == rank:
                for key in r1:
                    r[key][sidx - rank * S_loc] = r1[key][0]
            rescored[0] += 1

        def resolve_flagged(n, r):
            """world > 1: one host collective per run; every flagged (step, sample) is re-scored by all ranks together."""
            mask = torch.zeros((n, S), dtype=torch.int32)
            for k, sidx in pending:
                mask[k, sidx] = 1
            del pending[:]
            dist.all_reduce(mask, group=host_pg)
            for k, sidx in zip(*np.nonzero(mask.numpy())):
                rescore(int(sidx), r if int(k) == n - 1 else None)

        host_s = {"upload": 0.0, "launch": 0.0, "wait": 0.0}

        def timed_call(key, fn, *a):
            t_ = time.perf_counter()
            r_ = fn(*a)
            host_s[key] += time.perf_counter() - t_
            return r_

        def run_pipeline(n):
            """n complete steps, each from the upload of its samples to its results on the host."""
            for key in host_s:
                host_s[key] = 0.0
            for j in range(min(E2E_DEPTH, n)):
                timed_call("upload", up, pair[j])            # samples of steps 0 .. depth-1
            for j in range(min(E2E_DEPTH - 1, n)):
                timed_call("launch", launch, j)              # depth-1 steps queued on the GPU ahead of the host
            r = None
            for k in range(n):
                if k + E2E_DEPTH - 1 < n:
                    timed_call("launch", launch, k + E2E_DEPTH - 1)   # its samples were uploaded when step k-1 was finished
                r = timed_call("wait", finish, k)
                if k + E2E_DEPTH < n:
                    timed_call("upload", up, pair[k % E2E_DEPTH])     # H2D of step k+depth into the buffers step k has released
            if world > 1:
                resolve_flagged(n, r)
            return r

        run_pipeline(max(args.warmup, 1))
        barrier()
        rescored[0] = 0
        t0 = time.perf_counter()
        last = run_pipeline(args.steps)                      # includes filling the pipeline: the first upload overlaps nothing
        barrier()
        e2e_s = time.perf_counter() - t0
        for key in out:
            if key in last:
                out[key][...] = last[key]
        if use_grouped:
            h2d_bytes = gs.h2d_bytes
        else:
            h2d_bytes = h_off.nbytes + h_chr.nbytes + h_pos.nbytes + (h_idx.nbytes + h_tab.nbytes if coded is not None else h_wei.nbytes)
        clocks = sampler.stop() if rank == 0 else None

    m_sum = torch.tensor([float(out["m"].astype(np.int64).sum()), float((guard_resident > 0).sum())], dtype=torch.float64, device=dev)
    tms = torch.tensor([dev_ms, e2e_s * 1e3, hard_ms, exact_ms, hardg_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(m_sum)
    dev_ms, e2e_ms, hard_ms, exact_ms, hardg_ms = float(tms[0]), float(tms[1]), float(tms[2]), float(tms[3]), float(tms[4])

    if rank == 0:
        m_total = int(m_sum[0])
        comps = m_total * n_acc
        value = comps * args.steps / (dev_ms * 1e-3)
        e2e_value = comps * args.steps / (e2e_ms * 1e-3)
        # roofline of the scoring kernel on this rank: its share of the matched rows
        m_rank = m_total if world == 1 else None
        local_rows = None
        if world > 1:
            local_rows = sum(int(((s["rows"] >= r0) & (s["rows"] < r1)).sum()) for s in samples)
        rows_here = m_total if world == 1 else local_rows
        algo_bytes = rows_here * ((n_acc + 3) // 4 + 24) + 16 * n_acc * S
        k_ms = float(np.mean(score_ms))
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world, S),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "host_ms_per_step": {k: 1e3 * v / args.steps for k, v in host_s.items()},
                    "h2d_bytes_per_step": int(world * h2d_bytes),
                    "d2h_bytes_per_step": int(world * (sum(v.nbytes for v in out.values()) + 4 * S_loc)),
                    "inputs": "pinned host arrays in grouped order (snpm_group_markers, once at parse time: %.0f ms for the batch): "
                              "chromosome id and position in one uint32 per marker, weight-triple ids run-length coded (uint16 id + uint32 end per run; 4.1 bytes per marker) + the table of distinct triples "
                              "(f64); three batches rotate in a software pipeline (H2D of step k+3 and D2H of step k overlap the kernels "
                              "of steps k+1 and k+2; filling the pipeline is inside the timed region); the D2H holds scores, counts, "
                              "likelihoods and the per-sample guard counts" % (1e3 * t_group),
                    "samples_rescored_in_reference_order": int(rescored[0])},
            "gpu_launches": int(total_launches * args.steps),
            "roofline": {"bound": "hbm", "kernel": "k_score_grouped" if use_grouped else "k_score_segments", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, `ncu --set full` capture of this command
                         # (profiles/r1c_score_grouped_ncu_full.txt); only valid for the default single-GPU workload
                         "traffic": 1205027000 if (use_grouped and world == 1 and S == 64 and n_rows == N_ROWS and n_acc == N_ACC
                                                   and args.markers == N_DB_MARKERS) else None,
                         "algorithmic_bytes_per_launch": int(algo_bytes), "kernel_ms": k_ms,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "stages_ms": {k: float(np.mean(v)) for k, v in stage_ms.items() if k.endswith("_ms")},
            "distinct_weight_triples": int(len(gs.table)), "group_chunk_rows": int(args.group_chunk),
            "headline_kernel": "grouped counting kernel" if use_grouped else "order-exact fp64 kernel (row sharding leaves %.1f rows per weight group and rank)" % rows_per_group,
            "guard_flagged_samples": int(m_sum[1]),
            "clocks": clocks,
            "matched_markers_per_step": m_total,
        }
        # the same samples as called genotypes (0/1 weights): popcount kernel, bound by the HBM row gather
        hk_ms = float(np.mean(hard_kernel_ms))
        h_bytes = rows_here * ((n_acc + 3) // 4 + 5) + 12 * ((n_acc + 63) // 64 * 64) * int(np.ceil(args.markers / 1000.0)) * S
        h_ach = h_bytes / (hk_ms * 1e-3) / 1e9 if hk_ms > 0 else 0.0
        hgk_ms = float(np.mean(hardg_kernel_ms))
        hg_ach = h_bytes / (hgk_ms * 1e-3) / 1e9 if hgk_ms > 0 else 0.0
        line["called_genotypes"] = {
            "workload": "same batch with one-hot weights (BED / GT-only VCF inputs): grouped counting kernel (three weight groups per sample)",
            "value": comps * args.steps / (hardg_ms * 1e-3), "unit": UNIT, "ms_per_step": hardg_ms / args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_score_grouped", "achieved": hg_ach, "peak": peak, "unit": "GB/s",
                         "frac": hg_ach / peak if peak else None, "algorithmic_bytes_per_launch": int(h_bytes), "kernel_ms": hgk_ms},
            "popcount_kernel": {"kernel": "k_score_hard (position order, no host preparation)", "value": comps * args.steps / (hard_ms * 1e-3),
                                "ms_per_step": hard_ms / args.steps, "kernel_ms": hk_ms, "frac": h_ach / peak if peak else None,
                                "identical_results": bool(hard_same)}}
        xk_ms = float(np.mean(exact_kernel_ms))
        x_ach = algo_bytes / (xk_ms * 1e-3) / 1e9 if xk_ms > 0 else 0.0
        ok = guard_resident == 0
        line["order_exact_fp64"] = {
            "workload": "same batch in position order, fp64 kernel k_score_segments (sums in the reference's order: fp64 scores bit-identical)",
            "value": comps * args.steps / (exact_ms * 1e-3), "unit": UNIT, "ms_per_step": exact_ms / args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_score_segments", "achieved": x_ach, "peak": peak, "unit": "GB/s",
                         "frac": x_ach / peak if peak else None, "kernel_ms": xk_ms},
            "grouped_vs_exact": {"matches_equal": bool(np.array_equal(out["matches"][ok], exact_res["matches"][ok])),
                                 "ninfo_equal": bool(np.array_equal(out["ninfo"], exact_res["ninfo"])),
                                 "score_max_rel_diff": float(np.max(np.abs(out["score"] - exact_res["score"]) / np.maximum(exact_res["score"], 1.0))),
                                 "LR_max_rel_diff": float(np.nanmax(np.abs(out["LR"][ok] - exact_res["LR"][ok]) / np.abs(exact_res["LR"][ok])))}}
        if not args.no_cpu_baseline and world == 1:
            s0 = samples[0]
            rows, codes = cpu_prepare_sample(s0, n_acc, args.cpu_markers)
            c, dt, cpu_score, cpu_ninfo = cpu_run_sample(positions, regions, s0, rows, codes)
            line["cpu_baseline"] = {"value": c / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
                                    "sample": "sample 0 of the batch (%d matched markers x %d accessions): join over the %d "
                                              "database labels + chunked matchGTsAccs + likelihoods, NumPy oracle port of "
                                              "snpmatch.py:207-233, database rows held in RAM as int8" % (len(rows), n_acc, n_rows)}
            if not args.cpu_markers:
                # integers bit-exact (matches = int(score), informative sites); fp64 scores of the grouped kernel to 1e-12
                line["cpu_baseline"]["parity"] = bool(np.array_equal(cpu_score.astype(np.int64), out["matches"][0]) and
                                                      np.array_equal(cpu_ninfo, out["ninfo"][0]) and
                                                      np.allclose(cpu_score, out["score"][0], rtol=1e-12, atol=0.0) and
                                                      np.array_equal(cpu_score, exact_res["score"][0]))
        if world == 1:
            # batched shared-panel mode (BASELINE configs[3]): 4096 called-genotype samples on 20 000 shared markers as a
            # one-hot int8 GEMM on tcgen05; device time of the GEMM kernel (operand expansion and H2D of the codes excluded)
            rng = np.random.default_rng(9)
            S9, K9 = 4096, 20000
            rows9 = np.sort(rng.choice(n_rows, size=K9, replace=False))
            codes9 = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(S9, K9), p=[0.6, 0.28, 0.02, 0.1])
            g9 = min(db.score_shared_panel(rows9, codes9, likelihoods=False)["gemm_ms"] for _ in range(3))
            ops9 = 2.0 * (2 * S9) * n_acc * (3 * K9)          # algorithmic int8 ops (the kernel pads A to 1280 and K-slots to 4 per row)
            peak9 = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
            line["batched_shared_panel"] = {
                "workload": "configs[3]: %d called-genotype samples x %d shared markers vs %d accessions, one-hot int8 GEMM on tcgen05" % (S9, K9, n_acc),
                "value": S9 * K9 * n_acc / (g9 * 1e-3), "unit": UNIT, "gemm_ms": g9,
                "roofline": {"bound": "tensor", "kernel": "k_onehot_gemm", "achieved": ops9 / (g9 * 1e-3) / 1e12, "peak": peak9, "unit": "TOP/s (int8)",
                             "frac": ops9 / (g9 * 1e-3) / 1e12 / peak9,
                             "peak_source": "2 x measured dense bf16 burst TFLOP/s of MEASURED_PEAKS.json (int8 dense is nominally 2x bf16: 4500 vs 2250)"}}
        print(json.dumps(line))
    batch.close()
    for bx in extra:
        bx.close()
    gbatch.close()
    g.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
