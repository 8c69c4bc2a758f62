"""Parity of the coded path — position-order markers + weight codes uploaded as a parser hands them over, grouped by weight
triple ON THE DEVICE (csrc/group_sort.cuh: key build, stable segmented radix sort, change masks) and scored by the persistent
counting kernel k_score_grouped2 — against the CPU oracle and the order-exact fp64 kernel.

Bars as for the grouped kernel (tests/test_gpu_grouped.py): matches = int(score), ninfo, matched pairs bit-exact; fp64
scores rtol 1e-12; probabilities bit-exact; likelihoods and ratios rtol 1e-9 (north star asks 1e-6)."""
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


def _host_keys(cs):
    """The sort key of csrc/group_sort.cuh for every marker, in NumPy: called class | its code | slow class | fast class."""
    b = max(1, int(np.ceil(np.log2(max(len(cs.wtable), 2)))))
    cd = np.stack([cs.codes[:, 0], cs.codes[:, 2], cs.codes[:, 1]], axis=1).astype(np.int64)      # classes (ref, alt, het)
    w = cs.wtable[cd]
    c = np.where(w[:, 0] == 1.0, 0, np.where(w[:, 1] == 1.0, 1, np.where(w[:, 2] == 1.0, 2, -1)))
    big = np.zeros(len(w), dtype=np.int64)
    big = np.where(w[:, 1] > w[:, 0], 1, big)
    big = np.where(w[:, 2] > np.where(big == 1, w[:, 1], w[:, 0]), 2, big)
    c = np.where(c < 0, big, c)
    slow = np.where(c == 2, 0, 2)
    fast = np.where(c == 1, 0, 1)
    ix = np.arange(len(w))
    return (c << (3 * b)) | (cd[ix, c] << (2 * b)) | (cd[ix, slow] << b) | cd[ix, fast]


@pytest.mark.parametrize("n_acc", [1135, 33, 2100])
@pytest.mark.parametrize("skip", [False, True])
@pytest.mark.parametrize("group_kernel", ["sample", "tiled"])
def test_coded_path_vs_oracle_and_exact_kernel(lib, n_acc, skip, group_kernel, monkeypatch):
    # both forms of the device grouping: one kernel per sample (counters in shared memory) and the kernels over tiles of pairs
    monkeypatch.setenv("SNPM_GROUP_KERNEL", group_kernel)
    n_rows = 70000
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    sizes = [(2500, 200), (0, 50), (999, 0), (1000, 1), (1001, 7), (4321, 300), (1, 0), (17, 3), (20000, 1000)]
    samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=(3 + 11 * i) % n_acc, n_db=nd, n_extra=ne,
                                 seed=1900 + i, het=0.05) for i, (nd, ne) in enumerate(sizes)]
    offs, chrom, p, wei = _concat(samples)
    b = lib.Batch(db, offs, chrom, p, wei)
    b.run(skip_db_hets=skip)
    b.epilogue()
    exact = {k: v.copy() for k, v in b.fetch().items()}
    pairs = [b.fetch_pairs(i) for i in range(len(samples))]
    cs = lib.code_markers(offs, chrom, p, wei)
    assert cs is not None and len(cs.wtable) < 1024          # 32-bit keys
    keys = _host_keys(cs)
    for chunk in (320, 16, 208, 496):
        b.set_group_chunk(chunk)
        b.upload_coded(cs)
        b.run(skip_db_hets=skip, kernel_mode=lib.KERNEL_GROUPED)
        b.epilogue()
        r = b.fetch()
        guard = b.guard_counts()
        for i, s in enumerate(samples):
            if guard[i]:
                continue                # int(score) is decided by the reference's rounding: such samples are re-scored
            _check_against(r, {k: exact[k][i] for k in exact}, i)
        # the device grouping: pairs of a sample ordered by (key, position), the same set as the join's
        for i in (0, 5, 8):
            db_idx, s_idx = b.fetch_pairs(i)
            lo = int(offs[i])
            want = np.lexsort((pairs[i][1], keys[lo + pairs[i][1]]))
            assert np.array_equal(s_idx, pairs[i][1][want]) and np.array_equal(db_idx, pairs[i][0][want])
    assert guard.sum() <= 1
    for i in (0, 5, 8):
        ref_s, ref_n = _oracle_sample(pairs[i][0], pairs[i][1], samples[i]["wei"], n_acc, skip)
        lik, lr = orc.calculate_likelihoods(ref_s.astype(np.int64), ref_n)
        assert np.array_equal(r["matches"][i], ref_s.astype(np.int64)) and np.array_equal(r["ninfo"][i], ref_n)
        np.testing.assert_allclose(r["score"][i], ref_s, rtol=SCORE_RTOL)
        np.testing.assert_allclose(r["L"][i], lik, rtol=RTOL, equal_nan=True)
        np.testing.assert_allclose(r["LR"][i], lr, rtol=RTOL, equal_nan=True)
    t = b.coded_timings()
    assert t["score_ms"] > 0 and t["group_ms"] > 0
    b.close()
    db.close()


def test_coded_wide_panel_20000_accessions(lib):
    """BASELINE configs[4] width: 20 000 accessions (626 words per row, 18 word slices per segment), a short panel."""
    n_rows, n_acc = 4000, 20000
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=11 + 977 * i, n_db=nd, n_extra=ne, seed=7100 + i, het=0.05)
               for i, (nd, ne) in enumerate([(700, 30), (333, 5), (1, 0)])]
    offs, chrom, p, wei = _concat(samples)
    b = lib.Batch(db, offs, chrom, p, wei)
    b.run()
    b.epilogue()
    exact = {k: v.copy() for k, v in b.fetch().items()}
    pairs = b.fetch_pairs(0)
    cs = lib.code_markers(offs, chrom, p, wei)
    r = lib.score_coded(db, cs, chrom, p, wei, batch=b)
    for i in range(3):
        _check_against(r, {k: exact[k][i] for k in exact}, i, exact_scores=(i in set(r["rescored"].tolist())))
    ref_s, ref_n = _oracle_sample(pairs[0], pairs[1], samples[0]["wei"], n_acc)
    assert np.array_equal(r["matches"][0], ref_s.astype(np.int64)) and np.array_equal(r["ninfo"][0], ref_n)
    assert int(np.nanargmin(r["L"][0])) == 11
    # the host-grouped kernel of round 1 on the same width
    g = lib.group_markers(offs, chrom, p, wei)
    b.upload_grouped(g)
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    r1 = b.fetch()
    ok = b.guard_counts() == 0
    for i in np.flatnonzero(ok):
        _check_against(r1, {k: exact[k][i] for k in exact}, i)
    b.close()
    db.close()


def test_coded_called_genotypes_are_exact(lib):
    """One-hot weights: every score is an integer count and the result equals the popcount kernel bit for bit."""
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
    cs = lib.code_markers(offs, chrom, p, wei)
    assert len(cs.wtable) == 2
    b.upload_coded(cs)
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    r = b.fetch()
    assert b.guard_counts().sum() == 0
    for k in ("score", "matches", "ninfo", "m", "prob", "L", "LR"):
        assert np.array_equal(r[k], ref[k], equal_nan=True), k
    with pytest.raises(lib.SnpmError):
        b.run(kernel_mode=lib.KERNEL_FP64)                 # a coded batch is scored by the counting kernel only
    # codes outside the table are reported, not read
    bad = lib.CodedSamples(cs.offsets, cs.chrom_pos, np.full_like(cs.codes, 7), cs.wtable)
    b.upload_coded(bad)
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    with pytest.raises(lib.SnpmError):
        b.fetch()
    # chunks the persistent kernel cannot take are refused at upload
    b.set_group_chunk(1008)
    with pytest.raises(lib.SnpmError):
        b.upload_coded(cs)
    b.set_group_chunk(24)
    with pytest.raises(lib.SnpmError):
        b.upload_coded(cs)
    b.close()
    db.close()


def test_coded_many_distinct_weights_use_64_bit_keys(lib):
    """More than 1024 distinct weight values: 3 x 11 + 2 = 35 key bits, the 64-bit instantiation of sort and kernel; also
    triples with two exact ones, no one at all, zeros, values above one, and a group longer than the 240-row read-out limit."""
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
    for many in (False, True):
        wei = menu[pick].copy()
        if many:
            rare = rng.choice(n, size=1500, replace=False)
            wei[rare, 2] = rng.integers(1, 2000, size=1500) / 4096.0          # exact binary fractions: ~1300 extra distinct values
        offs = np.array([0, n])
        b = lib.Batch(db, offs, s["chr_ix"], s["pos"], wei)
        b.run()
        b.epilogue()
        exact = {k: v.copy() for k, v in b.fetch().items()}
        cs = lib.code_markers(offs, s["chr_ix"], s["pos"], wei)
        assert (len(cs.wtable) > 1024) == many
        r = lib.score_coded(db, cs, s["chr_ix"], s["pos"], wei, batch=b)
        _check_against(r, {k: exact[k][0] for k in exact}, 0, exact_scores=(0 in set(r["rescored"].tolist())))
        b.close()
    db.close()


def test_guard_band_catches_a_real_truncation_flip(lib):
    """Adversarial case for int(score) (snpmatch.py:96): ten markers of weight 0.1.  The reference adds 0.1 ten times and gets
    0.9999999999999999 -> 0 matches; counting gives fma(0.1, 10, 0) = 1.0 -> 1 match.  The guard must flag the cell and the
    re-scoring in reference order must restore 0 — for the coded path and for the host-grouped kernel of round 1."""
    n_acc = 40
    n_rows = 64
    pos = (np.arange(n_rows, dtype=np.int32) + 1) * 100
    regions = np.array([[0, n_rows]], dtype=np.int64)
    rng = np.random.default_rng(5)
    snps = rng.choice(np.array([0, 1, 2, -1], dtype=np.int8), size=(n_rows, n_acc), p=[0.5, 0.4, 0.05, 0.05])
    rows = np.arange(3, 33, 3)                            # the ten marker rows
    snps[rows, 0] = 0                                     # accession 0: ref everywhere -> ten times w_ref
    snps[rows, 1] = 1                                     # accession 1: alt everywhere
    snps[rows[:7], 2], snps[rows[7:], 2] = 0, 1           # accession 2: 7 ref + 3 alt
    db = lib.Database(pos, regions, n_acc)
    db.load_int8(snps)
    chrom = np.zeros(10, np.int32)
    p = pos[rows]
    wei = np.full((10, 3), 0.1)
    offs = np.array([0, 10])
    ref_s, ref_n = orc.match_gts_accs(wei, snps[rows].copy(), False)
    assert ref_s[0] == 0.9999999999999999 and int(ref_s[0]) == 0          # the reference's own rounding
    b = lib.Batch(db, offs, chrom, p, wei)
    b.run()
    b.epilogue()
    exact = {k: v.copy() for k, v in b.fetch().items()}
    assert np.array_equal(exact["score"][0], ref_s) and np.array_equal(exact["matches"][0], ref_s.astype(np.int64))
    cs = lib.code_markers(offs, chrom, p, wei)
    b.upload_coded(cs)
    b.run(kernel_mode=lib.KERNEL_GROUPED)
    b.epilogue()
    raw = b.fetch()
    assert b.guard_counts()[0] > 0
    assert raw["matches"][0][0] == 1 and exact["matches"][0][0] == 0      # the flip is real before re-scoring
    r = lib.score_coded(db, cs, chrom, p, wei, batch=b)
    assert r["rescored"].tolist() == [0]
    _check_against(r, {k: exact[k][0] for k in exact}, 0, exact_scores=True)
    # host-grouped kernel (round 1 path)
    r1 = lib.score_grouped(db, offs, chrom, p, wei, batch=b)
    assert r1["rescored"].tolist() == [0]
    _check_against(r1, {k: exact[k][0] for k in exact}, 0, exact_scores=True)
    b.close()
    db.close()


def test_coded_edge_cases(lib):
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
        cs = lib.code_markers(offs, chrom, p, wei)
        r = lib.score_coded(db, cs, chrom, p, wei, batch=b)
        assert r["m"].tolist() == exact["m"].tolist() and r["m"][0] == 0 and r["m"][2] == 0
        for i in range(5):
            _check_against(r, {k: exact[k][i] for k in exact}, i, exact_scores=(i in set(r["rescored"].tolist())))
        assert np.all(r["matches"][3] == 0) and np.all(np.isnan(r["L"][0])) and np.all(np.isnan(r["L"][3]))
        b.close()
        db.close()
    # an entirely empty batch
    db = lib.Database(pos, regions, 40)
    db.fill_synthetic(synth.SEED_PANEL)
    cs = lib.code_markers(np.array([0, 0]), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3)))
    r = lib.score_coded(db, cs)
    assert r["m"].tolist() == [0] and np.all(r["ninfo"] == 0)
    db.close()


def test_coded_full_shape_properties(lib):
    """BASELINE configs[1] shape (10.7 M x 1135 panel, 50 k-marker PL samples): size-independent properties — the coded path and
    the order-exact kernel agree on every integer for 4 samples, the true accessions are recovered, ninfo <= matched rows,
    a second run of the same batch reproduces the first bit for bit (the device grouping is deterministic)."""
    n_rows, n_acc = 10_700_000, 1135
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    samples = [synth.make_sample_fast(pos, regions, n_acc, true_acc=7 + 13 * i, n_db=45000, n_extra=5000, seed=5000 + i) for i in range(4)]
    offs, chrom, p, wei = _concat(samples)
    cs = lib.code_markers(offs, chrom, p, wei)
    b = lib.Batch(db, offs, chrom, p, wei)
    r = {k: v.copy() for k, v in lib.score_coded(db, cs, chrom, p, wei, batch=b).items()}
    r2 = lib.score_coded(db, cs, chrom, p, wei, batch=b)
    for k in ("score", "matches", "ninfo", "L", "LR"):
        assert np.array_equal(r[k], r2[k], equal_nan=True), k
    b.upload(offs, chrom, p, wei)
    b.run()
    b.epilogue()
    exact = b.fetch()
    for i in range(4):
        _check_against(r, {k: exact[k][i] for k in exact}, i, exact_scores=(i in set(r["rescored"].tolist())))
        assert int(np.nanargmin(r["L"][i])) == 7 + 13 * i
        assert r["ninfo"][i].max() <= r["m"][i] == 45000
    b.close()
    db.close()
