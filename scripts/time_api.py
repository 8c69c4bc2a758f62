"""Development helper: where the time of core.batch.genotype_many goes (host preparation / device call / flagged samples)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from snpmatch_b200 import lib, synth
from snpmatch_b200.core import batch, parsers, snp_genotype

rows, n_acc = 10_700_000, 1135
positions, regions = synth.panel_positions(rows)
g = snp_genotype.Genotype.synthetic(rows, n_acc, device=0)
bench.N_EXTRA_MARKERS = 5000
samples = bench.make_samples(positions, regions, n_acc, 64, 45000)
names = np.array(["Chr" + c for c in synth.TAIR10_CHRS])
inputs = []
for s in samples:
    inp = parsers.ParseInputs("")
    inp.load_snp_info(names[s["chr_ix"]], s["pos"], synth._gt_strings(s["code"]), s["wei"], s["dp"])
    inp._coded = (s["pl"].astype(np.uint16), synth.pl_table(int(s["pl"].max())), inp.wei)
    inputs.append(inp)
for rep in range(3):
    t0 = time.perf_counter()
    cs, offs, cid, pos, wei = batch.coded_batch(g, inputs, with_weights=False)
    t1 = time.perf_counter()
    r = lib.score_coded(g.db, cs, cid, pos, wei)
    t2 = time.perf_counter()
    res = batch.genotype_many(g, inputs)
    t3 = time.perf_counter()
    print("coded_batch %.1f ms  score_coded %.1f ms (rescored %d)  genotype_many %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), len(r["rescored"]), 1e3 * (t3 - t2)))
b = lib.Batch(g.db, [0, 0], np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3)))
for rep in range(3):
    t = [time.perf_counter()]
    b.upload_coded(cs); t.append(time.perf_counter())
    b.run(False, kernel_mode=lib.KERNEL_GROUPED); t.append(time.perf_counter())
    b.epilogue(); t.append(time.perf_counter())
    r = b.fetch(); t.append(time.perf_counter())
    gc = b.guard_counts(); t.append(time.perf_counter())
    print("upload %.2f run %.2f epilogue %.2f fetch %.2f guard %.2f ms; flagged %d" % tuple([1e3 * (t[i + 1] - t[i]) for i in range(5)] + [int((gc > 0).sum())]))
