"""Parity of the grouped counting kernel (k_score_grouped, kernel mode 2) — the throughput path of likelihood-weighted
scoring — against the CPU oracle and against the order-exact fp64 kernel.

Bars: matches = int(score), ninfo, matched pairs bit-exact; fp64 scores within rtol 1e-12 of the reference-order sum
(they are bit-exact when every weight is 0/1); probabilities bit-exact (ratios of exact integers); likelihoods and
ratios rtol 1e-9 (north star asks 1e-6)."""
import numpy as np
import pytest

from oracle import snpmatch_oracle as orc
from snpmatch_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-9
SCORE_RTOL = 1e-12


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from snpmatch_b200 import lib as L
    assert L.device_count() > 0, "GPU tests need a CUDA device"
    return L


def _concat(samples, wei_key="wei"):
    offs = np.concatenate([[0], np.cumsum([len(s["pos"]) for s in samples])])
    return (offs, np.concatenate([s["chr_ix"] for s in samples]), np.concatenate([s["pos"] for s in samples]),
            np.concatenate([s[wei_key] for s in samples]))


def _oracle_sample(db_idx, s_idx, wei, n_acc, skip=False):
    """Genotyper.genotyper's chunk loop (snpmatch.py:218-225) on the panel rows of the matched pairs."""
    codes = synth.panel_codes(synth.SEED_PANEL, db_idx, n_acc)
    score, ninfo = np.zeros(n_acc), np.zeros(n_acc, dtype=np.int64)
    for j in range(0, len(db_idx), 1000):
        t_s, t_n = orc.match_gts_accs(wei[s_idx[j:j + 1000]], codes[j:j + 1000].copy(), skip)
        score, ninfo = score + t_s, ninfo + t_n
    return score, ninfo


def _check_against(r, ref, i, exact_scores=False):
    assert np.array_equal(r["matches"][i], ref["matches"]), "matches differ"
    assert np.array_equal(r["ninfo"][i], ref["ninfo"]), "ninfo differs"
    assert int(r["m"][i]) == int(ref["m"])
    if exact_scores:
        assert np.array_equal(r["score"][i], ref["score"])
    else:
        np.testing.assert_allclose(r["score"][i], ref["score"], rtol=SCORE_RTOL, atol=0)
    np.testing.assert_array_equal(r["prob"][i], ref["prob"])
    np.testing.assert_allclose(r["L"][i], ref["L"], rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(r["LR"][i], ref["LR"], rtol=RTOL, equal_nan=True)


@pytest.mark.parametrize("n_acc", [1135, 33, 2100])
@pytest.mark.parametrize("skip", [False, True])
def test_grouped_kernel_vs_oracle_and_exact_kernel(lib, n_acc, skip):
    n_rows = 70000
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    sizes = [(2500, 200), (0, 50), (999, 0), (1000, 1), (1001, 7), (4321, 300), (1, 0), (17, 3), (20000, 1000)]
    samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=(3 + 11 * i) % n_acc, n_db=nd, n_extra=ne,
                                 seed=1900 + i, het=0.05) for i, (nd, ne) in enumerate(sizes)]
    offs, chrom, p, wei = _concat(samples)
    # order-exact kernel (bit-exact against the oracle, tests/test_gpu_parity.py)
    b = lib.Batch(db, offs, chrom, p, wei)
    b.run(skip_db_hets=skip)
    b.epilogue()
    exact = {k: v.copy() for k, v in b.fetch().items()}
    pairs = [b.fetch_pairs(i) for i in range(len(samples))]
    # grouped kernel
    g = lib.group_markers(offs, chrom, p, wei)
    assert g is not None and len(g.table) < 5000
    for chunk in (320, 16, 208, 1008):
        b.set_group_chunk(chunk)
        b.upload_grouped(g)
        b.run(skip_db_hets=skip, kernel_mode=lib.KERNEL_GROUPED)
        b.epilogue()
        r = b.fetch()
        guard = b.guard_counts()
        for i, s in enumerate(samples):
            if guard[i]:
                continue                # int(score) is decided by the reference's rounding: such samples are re-scored (below)
            _check_against(r, {k: exact[k][i] for k in exact}, i)
    assert guard.sum() <= 1
    # oracle, directly, for three samples
    for i in (0, 5, 8):
        ref_s, ref_n = _oracle_sample(pairs[i][0], pairs[i][1], samples[i]["wei"], n_acc, skip)
        lik, lr = orc.calculate_likelihoods(ref_s.astype(np.int64), ref_n)
        assert np.array_equal(r["matches"][i], ref_s.astype(np.int64)) and np.array_equal(r["ninfo"][i], ref_n)
        np.testing.assert_allclose(r["score"][i], ref_s, rtol=SCORE_RTOL)
        np.testing.assert_allclose(r["L"][i], lik, rtol=RTOL, equal_nan=True)
        np.testing.assert_allclose(r["LR"][i], lr, rtol=RTOL, equal_nan=True)
    if n_acc > 100:
        assert int(np.nanargmin(r["L"][8])) == (3 + 11 * 8) % n_acc
    # the matched pairs of a grouped batch are the same set
    db_idx, s_idx = b.fetch_pairs(5)
    assert np.array_equal(np.sort(db_idx), pairs[5][0])
    assert np.array_equal(pos[db_idx].astype(np.int64), g.pos[int(offs[5]):int(offs[6])][s_idx])
    b.close()
    db.close()


def test_grouped_called_genotypes_are_exact(lib):
    """One-hot weights: every score is an integer count, F stays 0 and the result equals the popcount kernel bit for bit."""
    n_rows, n_acc = 50000, 1135
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=5 + 7 * i, n_db=nd, n_extra=ne, seed=2700 + i, het=0.05)
               for i, (nd, ne) in enumerate([(3333, 100), (1, 0), (1000, 10), (2049, 5), (0, 3)])]
    offs, chrom, p, wei = _concat(samples, "wei_hard")
    b = lib.Batch(db, offs, chrom, p, wei)
    b.run(kernel_mode=lib.KERNEL_POPCOUNT)
    b.epilogue()
    ref = {k: v.copy() for k, v in b.fetch().items()}
    g = lib.group_markers(offs, chrom, p, wei)
    assert len(g.table) == 3
    b.upload_grouped(g)
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    r = b.fetch()
    assert b.guard_counts().sum() == 0
    for k in ("score", "matches", "ninfo", "m", "prob", "L", "LR"):
        assert np.array_equal(r[k], ref[k], equal_nan=True), k
    # the unpacked upload (chromosome byte + position word) gives the same
    assert g.packed is not None
    g.packed = None
    b.upload_grouped(g)
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    r2 = b.fetch()
    for k in ("score", "matches", "ninfo", "m"):
        assert np.array_equal(r2[k], ref[k]), k
    # a grouped batch refuses the other kernels, windows and the F1 pass
    with pytest.raises(lib.SnpmError):
        b.run(kernel_mode=lib.KERNEL_FP64)
    # weight-triple ids outside the table are reported, not read
    bad = lib.GroupedSamples(g.offsets, g.chrom, g.pos, np.full_like(g.gid, 7), g.table, g.order)
    b.upload_grouped(bad)
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    with pytest.raises(lib.SnpmError):
        b.fetch()
    b.close()
    db.close()


def test_grouped_unusual_weight_triples(lib):
    """Triples with two exact ones, no one at all, zeros, values above one, and a group longer than the 255-row counters."""
    n_rows, n_acc = 30000, 300
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    rng = np.random.default_rng(77)
    s = synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=9, n_db=6000, n_extra=100, seed=31, het=0.05)
    n = len(s["pos"])
    menu = np.array([[1.0, 1.0, 0.25], [0.3, 0.2, 0.7], [0.0, 0.0, 0.0], [2.5, 1.0, 0.0], [1.0, 0.0, 1e-30], [1e-300, 0.125, 1.0],
                     [0.9999999999999999, 1.0000000000000002, 1.0]])
    pick = rng.choice(len(menu), size=n, p=[0.05, 0.6, 0.05, 0.05, 0.1, 0.05, 0.1])
    wei = menu[pick]
    offs = np.array([0, n])
    b = lib.Batch(db, offs, s["chr_ix"], s["pos"], wei)
    b.run()
    b.epilogue()
    exact = {k: v.copy() for k, v in b.fetch().items()}
    r = lib.score_grouped(db, offs, s["chr_ix"], s["pos"], wei, batch=b)
    _check_against(r, {k: exact[k][0] for k in exact}, 0)
    b.close()
    db.close()


def test_grouped_guard_band_triggers_rescoring(lib):
    """Weights of exactly 0.5 make F an exact integer for every accession with an even count: the guard cannot tell that
    from a rounding accident, flags the sample, and score_grouped re-scores it with the order-exact kernel."""
    n_rows, n_acc = 20000, 200
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    s0 = synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=4, n_db=3000, n_extra=10, seed=41)
    s1 = synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=8, n_db=2000, n_extra=10, seed=42)
    w1 = np.where(s1["wei_hard"] == 1.0, 1.0, 0.5)
    offs = np.array([0, len(s0["pos"]), len(s0["pos"]) + len(s1["pos"])])
    chrom = np.concatenate([s0["chr_ix"], s1["chr_ix"]])
    p = np.concatenate([s0["pos"], s1["pos"]])
    wei = np.concatenate([s0["wei"], w1])
    b = lib.Batch(db, offs, chrom, p, wei)
    b.run()
    b.epilogue()
    exact = {k: v.copy() for k, v in b.fetch().items()}
    g = lib.group_markers(offs, chrom, p, wei)
    b.upload_grouped(g)
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    raw = b.fetch()
    guard = b.guard_counts()
    assert guard[1] > 0
    # even unre-scored, the grouped integers are right here (0.5 sums are exact in any order)
    assert np.array_equal(raw["matches"][1], exact["matches"][1])
    r = lib.score_grouped(db, offs, chrom, p, wei, batch=b)
    assert 1 in set(r["rescored"].tolist())
    for i in (0, 1):
        _check_against(r, {k: exact[k][i] for k in exact}, i, exact_scores=(i in set(r["rescored"].tolist())))
    b.close()
    db.close()


def test_grouped_full_shape_properties(lib):
    """BASELINE configs[1] shape (10.7 M x 1135 panel, 50 k-marker PL samples): size-independent properties — the grouped and
    the order-exact kernels agree on every integer for 4 samples, the true accessions are recovered, ninfo <= matched rows."""
    n_rows, n_acc = 10_700_000, 1135
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    samples = [synth.make_sample_fast(pos, regions, n_acc, true_acc=7 + 13 * i, n_db=45000, n_extra=5000, seed=5000 + i) for i in range(4)]
    offs, chrom, p, wei = _concat(samples)
    r = lib.score_grouped(db, offs, chrom, p, wei)
    b = lib.Batch(db, offs, chrom, p, wei)
    b.run()
    b.epilogue()
    exact = b.fetch()
    for i in range(4):
        _check_against(r, {k: exact[k][i] for k in exact}, i, exact_scores=(i in set(r["rescored"].tolist())))
        assert int(np.nanargmin(r["L"][i])) == 7 + 13 * i
        assert r["ninfo"][i].max() <= r["m"][i] == 45000
    b.close()
    db.close()


def test_async_fetch_overlaps_two_batches(lib):
    """snpm_batch_fetch_async / _wait: the read-back of one batch is queued, another batch runs, both results are right."""
    n_rows, n_acc = 40000, 500
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    sets = []
    for j in range(2):
        samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=2 + 5 * i + j, n_db=1500 + 700 * i, n_extra=50,
                                     seed=4000 + 10 * j + i) for i in range(3)]
        sets.append(_concat(samples))
    ref = []
    for offs, chrom, p, wei in sets:
        b = lib.Batch(db, offs, chrom, p, wei)
        b.run()
        b.epilogue()
        ref.append({k: v.copy() for k, v in b.fetch().items()})
        b.close()
    batches, outs = [], []
    for offs, chrom, p, wei in sets:
        b = lib.Batch(db, offs, chrom, p, wei)
        b.upload_grouped(lib.group_markers(offs, chrom, p, wei))
        S = len(offs) - 1
        out = {k: np.empty((S, n_acc), np.float64) for k in ("score", "prob", "L", "LR")}
        out.update({k: np.empty((S, n_acc), np.int64) for k in ("matches", "ninfo")})
        out["m"] = np.empty(S, np.int64)
        out["guard"] = np.full(S, -1, np.int32)
        batches.append(b)
        outs.append(out)
    with pytest.raises(lib.SnpmError):
        batches[0].fetch_wait()                      # nothing pending yet
    for b, out in zip(batches, outs):                # both queued before either is waited for
        b.run(kernel_mode=lib.KERNEL_GROUPED)
        b.epilogue()
        b.fetch_async(out)
    for j in (1, 0):
        r = batches[j].fetch_wait()
        ok = r["guard"] == 0
        assert ok.all()
        for i in range(len(r["m"])):
            _check_against(r, {k: ref[j][k][i] for k in ref[j]}, i)
    for b in batches:
        b.close()
    db.close()


@pytest.mark.parametrize("grouped", [False, True])
def test_result_range_restricts_epilogue_and_fetch(lib, grouped):
    """snpm_batch_set_result_range: the share of the samples one rank finishes after a reduce-scatter."""
    n_rows, n_acc = 30000, 257
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=1 + 3 * i, n_db=900 + 300 * i, n_extra=20, seed=5100 + i)
               for i in range(4)]
    offs, chrom, p, wei = _concat(samples)
    b = lib.Batch(db, offs, chrom, p, wei)
    mode = lib.KERNEL_GROUPED if grouped else lib.KERNEL_FP64
    if grouped:
        b.upload_grouped(lib.group_markers(offs, chrom, p, wei))
    b.run(kernel_mode=mode)
    b.epilogue()
    full = {k: v.copy() for k, v in b.fetch().items()}
    b.set_result_range(1, 2)
    b.run(kernel_mode=mode)
    b.epilogue()
    part = b.fetch()
    assert part["score"].shape == (2, n_acc) and len(part["m"]) == 2 and len(b.guard_counts()) == 2
    for k in full:
        assert np.array_equal(part[k], full[k][1:3], equal_nan=True), k
    b.set_result_range(3, 2)                     # past the end
    b.run(kernel_mode=mode)
    with pytest.raises(lib.SnpmError):
        b.epilogue()
    b.set_result_range()                         # back to all samples
    b.run(kernel_mode=mode)
    b.epilogue()
    again = b.fetch()
    for k in full:
        assert np.array_equal(again[k], full[k], equal_nan=True), k
    b.close()
    db.close()


def test_genotype_many_equals_per_sample_genotyper(lib, tmp_path):
    """Python mirror: core.batch.genotype_many (grouped kernel, PL and called samples mixed) vs Genotyper sample by sample."""
    from conftest import load_golden
    from snpmatch_b200.core import batch, parsers, snp_genotype, snpmatch
    p = load_golden("small_panel.npz")
    g = snp_genotype.Genotype.from_arrays(p["snps"], p["positions"], p["chrs"], p["chr_regions"], p["accessions"])
    inputs = []
    for i in range(8):
        s = synth.make_sample(p["positions"], p["chr_regions"], p["chrs"], 40, true_acc=3 + 4 * i, n_db=700 + 90 * i, n_extra=40, seed=600 + i)
        inp = parsers.ParseInputs("")
        inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s["wei_hard"] if i % 3 == 2 else s["wei"], s["dp"])
        inputs.append(inp)
    for skip in (False, True):
        results = batch.genotype_many(g, inputs, skip_db_hets=skip)
        for i, inp in enumerate(inputs):
            one = snpmatch.Genotyper(inp, g, str(tmp_path / ("m%d" % i)), run_genotyper=False, skip_db_hets=skip).genotyper()
            got = results[i]
            assert np.array_equal(got.scores, one.scores) and np.array_equal(got.ninfo, one.ninfo)
            assert got.num_snps == one.num_snps and got.overlap == one.overlap
            got.get_likelihoods(); one.get_likelihoods()
            np.testing.assert_allclose(got.likelis, one.likelis, rtol=RTOL, equal_nan=True)
            np.testing.assert_allclose(got.lrts, one.lrts, rtol=RTOL, equal_nan=True)
    g.close()


def test_grouped_edge_cases(lib):
    """Empty samples, a sample without any panel marker, all-zero weights, a one-accession panel, one marker."""
    n_rows = 5000
    pos, regions = synth.panel_positions(n_rows)
    for n_acc in (1, 40):
        db = lib.Database(pos, regions, n_acc)
        db.fill_synthetic(synth.SEED_PANEL)
        s = synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=0, n_db=300, n_extra=20, seed=77)
        n = len(s["pos"])
        miss_pos = np.setdiff1d(np.arange(1, 4000, dtype=np.int64), pos[:regions[0][1]].astype(np.int64))[:50]      # chromosome 0, not in the panel
        offs = np.array([0, 0, n, n + 50, n + 50 + n, n + 51 + n])
        chrom = np.concatenate([s["chr_ix"], np.zeros(50, np.int32), s["chr_ix"], s["chr_ix"][:1]])
        p = np.concatenate([s["pos"], miss_pos, s["pos"], s["pos"][:1]])
        wei = np.concatenate([s["wei"], np.full((50, 3), 0.5), np.zeros((n, 3)), s["wei"][:1]])
        b = lib.Batch(db, offs, chrom, p, wei)
        b.run()
        b.epilogue()
        exact = {k: v.copy() for k, v in b.fetch().items()}
        r = lib.score_grouped(db, offs, chrom, p, wei, batch=b)
        assert r["m"].tolist() == exact["m"].tolist() and r["m"][0] == 0 and r["m"][2] == 0
        for i in range(5):
            _check_against(r, {k: exact[k][i] for k in exact}, i, exact_scores=(i in set(r["rescored"].tolist())))
        assert np.all(r["matches"][3] == 0) and np.all(np.isnan(r["L"][0])) and np.all(np.isnan(r["L"][3]))
        b.close()
        db.close()


def test_grouped_upload_forms_agree(lib):
    """The three wire forms of a grouped batch (7, 6 and ~4 bytes per marker: separate chromosome/position, packed word,
    packed word + run-length coded weight-triple ids) give identical results; malformed runs are refused."""
    n_rows, n_acc = 60000, 257
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=5 + i, n_db=nd, n_extra=ne, seed=2300 + i)
               for i, (nd, ne) in enumerate([(3000, 100), (0, 10), (1, 0), (777, 33)])]
    offs, chrom, p, wei = _concat(samples)
    g = lib.group_markers(offs, chrom, p, wei)
    assert g.packed is not None and g.run_gid is not None and int(g.run_end[-1]) == len(p)
    assert g.h2d_bytes < offs.nbytes + 6 * len(p) + len(g.table) * 32
    b = lib.Batch(db, offs, chrom, p, wei)
    res = []
    for form in ("runs", "packed", "plain"):
        h = lib.GroupedSamples(g.offsets, g.chrom, g.pos, g.gid, g.table, g.order,
                               packed=None if form == "plain" else g.packed,
                               run_gid=g.run_gid if form == "runs" else None, run_end=g.run_end if form == "runs" else None)
        b.upload_grouped(h)
        b.run(kernel_mode=lib.KERNEL_GROUPED)
        b.epilogue()
        res.append({k: v.copy() for k, v in b.fetch().items()})
    for r in res[1:]:
        for k in res[0]:
            assert np.array_equal(r[k], res[0][k], equal_nan=True), k
    bad_end = g.run_end.copy()
    bad_end[-1] -= 1
    with pytest.raises(lib.SnpmError):
        b.upload_grouped(lib.GroupedSamples(g.offsets, g.chrom, g.pos, g.gid, g.table, g.order, packed=g.packed, run_gid=g.run_gid, run_end=bad_end))
    assert len(g.run_end) > 3
    bad_end = g.run_end.copy()
    bad_end[2] = bad_end[1]                    # an interior run that does not ascend: counted on the device, reported at fetch
    b.upload_grouped(lib.GroupedSamples(g.offsets, g.chrom, g.pos, g.gid, g.table, g.order, packed=g.packed, run_gid=g.run_gid, run_end=bad_end))
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    with pytest.raises(lib.SnpmError):
        b.fetch()
    b.close()
    db.close()
