"""Parity of the CUDA path (through the C ABI / the Python mirror of the reference interface) against
the CPU oracle and the golden vectors generated from the unmodified reference.

Bars: integers, indices and — because the kernels sum in the reference's order — the fp64 scores are
bit-exact; likelihoods / ratios / probabilities agree to rtol 1e-9 (north star asks 1e-6)."""
import json
import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import snpmatch_oracle as orc
from snpmatch_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from snpmatch_b200 import lib as L
    assert L.device_count() > 0, "GPU tests need a CUDA device"
    return L


@pytest.fixture(scope="module")
def small_geno(lib, small_panel):
    from snpmatch_b200.core import snp_genotype
    p = small_panel
    g = snp_genotype.Genotype.from_arrays(p["snps"], p["positions"], p["chrs"], p["chr_regions"], p["accessions"])
    yield g
    g.close()


# ---- A0 -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_acc", [1, 31, 32, 33, 64, 70, 300, 1135])
def test_pack_roundtrip_and_layout(lib, n_acc):
    rng = np.random.default_rng(n_acc)
    n = 257
    snps = rng.choice(np.array([-1, 0, 1, 2], dtype=np.int8), size=(n, n_acc), p=[0.1, 0.5, 0.3, 0.1])
    pos = np.arange(1, n + 1, dtype=np.int32)
    db = lib.Database(pos, np.array([[0, n]]), n_acc)
    db.load_int8(snps)
    assert np.array_equal(db.read_packed(0, n), orc.pack_2bit_words(snps))
    rows = rng.integers(0, n, size=100)
    assert np.array_equal(db.read_rows(rows), snps[rows])
    # packed load path
    db2 = lib.Database(pos, np.array([[0, n]]), n_acc)
    db2.load_packed(orc.pack_2bit_words(snps))
    assert np.array_equal(db2.read_rows(np.arange(n)), snps)
    db.close(); db2.close()


def test_synthetic_fill_matches_host_hash(lib):
    n_rows, n_acc = 5000, 333
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    rows = np.array([0, 1, 2, 77, 1234, 4998, 4999])
    assert np.array_equal(db.read_rows(rows), synth.panel_codes(synth.SEED_PANEL, rows, n_acc))
    db.close()
    # a shard generates the same global rows
    db = lib.Database(pos[1000:3000], np.clip(regions, 1000, 3000) - 1000, n_acc, row0_global=1000)
    db.fill_synthetic(synth.SEED_PANEL)
    assert np.array_equal(db.read_rows(np.array([0, 5, 1999])), synth.panel_codes(synth.SEED_PANEL, np.array([1000, 1005, 2999]), n_acc))
    db.close()


# ---- A2 -------------------------------------------------------------------------------------------
def test_match_gts_accs_golden_bit_exact(lib):
    from snpmatch_b200.core import snpmatch
    g = load_golden("match_gts_accs.npz")
    for i in range(int(g["n_cases"])):
        score, ninfo = snpmatch.matchGTsAccs(g["c%d_wei" % i], g["c%d_snps" % i].copy(), bool(g["c%d_skip" % i]))
        assert np.array_equal(ninfo, g["c%d_ninfo" % i]), "ninfo, case %d" % i
        assert np.array_equal(score, g["c%d_score" % i]), "fp64 score not bit-exact, case %d" % i


@pytest.mark.parametrize("k,n_acc", [(0, 5), (1, 1), (63, 40), (64, 129), (65, 1135), (1000, 1135), (2500, 777), (130, 2000)])
def test_match_gts_accs_vs_oracle(lib, k, n_acc):
    rng = np.random.default_rng(k * 7 + n_acc)
    snps = rng.choice(np.array([-1, 0, 1, 2], dtype=np.int8), size=(k, n_acc), p=[0.1, 0.55, 0.3, 0.05])
    code = rng.choice([0, 1, 2], size=k, p=[0.7, 0.25, 0.05]).astype(np.int8)
    wei = synth._pl_weights(rng, code, 1 + rng.poisson(3, size=k))[1] if k else np.zeros((0, 3))
    for skip in (False, True):
        score, ninfo = lib.match_gts_accs(wei, snps, skip)
        ref_s, ref_n = orc.match_gts_accs(wei, snps, skip)
        assert np.array_equal(ninfo, ref_n)
        assert np.array_equal(score, ref_s)


def test_match_gts_accs_edge_weights(lib):
    # many exact 1.0 weights plus tiny ones: the truncation edge of SURVEY section 7 hard part 1
    k, n_acc = 3000, 64
    rng = np.random.default_rng(1)
    snps = rng.choice(np.array([0, 1], dtype=np.int8), size=(k, n_acc))
    wei = np.zeros((k, 3))
    wei[:, 0] = 1.0
    wei[:, 2] = np.exp(-rng.integers(200, 900, size=k) / 10.0)
    snps[:, 5] = 0
    score, ninfo = lib.match_gts_accs(wei, snps)
    ref_s, ref_n = orc.match_gts_accs(wei, snps)
    assert np.array_equal(score, ref_s) and np.array_equal(ninfo, ref_n)
    assert score[5] == 3000.0 and int(score[5]) == 3000


# ---- A4 -------------------------------------------------------------------------------------------
def test_likelihood_known_answers(lib):
    from snpmatch_b200.core import snpmatch
    assert abs(snpmatch.likeliTest(10, 3) - 122.8361221819443) <= 1e-12 * 122.8361221819443   # tests/test_inbred.py:22
    assert np.isnan(snpmatch.likeliTest(10, 0))                                               # tests/test_inbred.py:24
    assert np.isnan(snpmatch.likeliTest(0, 0))
    assert snpmatch.likeliTest(7, 7) == 1.0
    with pytest.raises(AssertionError):
        snpmatch.likeliTest(3, 4)


def test_epilogue_golden_sets(lib):
    from snpmatch_b200.core import snpmatch
    g = load_golden("epilogue.npz")
    for i in range(int(g["n_sets"])):
        L, LR = snpmatch.GenotyperOutput.calculate_likelihoods(g["e%d_y" % i], g["e%d_n" % i])
        np.testing.assert_allclose(L, g["e%d_L" % i], rtol=RTOL, equal_nan=True)
        np.testing.assert_allclose(LR, g["e%d_LR" % i], rtol=RTOL, equal_nan=True)
    prob, L, LR = lib.calculate_likelihoods(np.array([3.0, 5.0]), np.array([10.0, 10.0]), amin=2.0)
    np.testing.assert_allclose(LR, L / 2.0, rtol=1e-15)
    np.testing.assert_allclose(prob, [0.3, 0.5], rtol=0)


# ---- A1 -------------------------------------------------------------------------------------------
def test_join_golden_cases(lib):
    from snpmatch_b200.core import snp_genotype
    g = load_golden("join.npz")
    for i in range(int(g["n_cases"])):
        i1, i2 = snp_genotype.Genotype.get_common_positions(g["j%d_c1" % i], g["j%d_p1" % i], g["j%d_c2" % i], g["j%d_p2" % i])
        assert np.array_equal(i1, g["j%d_i1" % i]), "db side, case %d" % i
        assert np.array_equal(i2, g["j%d_i2" % i]), "sample side, case %d" % i


@pytest.mark.parametrize("algo", [1, 2])
@pytest.mark.parametrize("n_db,n_s", [(6000, 0), (6000, 1), (6000, 2900), (200000, 5000), (200000, 150000)])
def test_join_search_and_mergepath_agree_with_oracle(lib, algo, n_db, n_s):
    rng = np.random.default_rng(n_db + n_s + algo)
    pos, regions = synth.panel_positions(n_db)
    db = lib.Database(pos, regions, 3)
    rows = np.sort(rng.choice(n_db, size=min(n_s, n_db) * 9 // 10, replace=False))
    chrom = np.searchsorted(regions[:, 1], rows, side="right")
    extra = n_s - len(rows)
    e_chr = rng.integers(0, 6, size=extra)                    # chromosome 5 does not exist in the panel
    e_pos = rng.integers(1, 18_000_000, size=extra)
    s_chr = np.concatenate([chrom, e_chr])
    s_pos = np.concatenate([pos[rows].astype(np.int64), e_pos])
    key = np.unique(s_chr.astype(np.int64) * (1 << 32) + s_pos)
    s_chr, s_pos = (key >> 32).astype(np.int32), (key & 0xFFFFFFFF).astype(np.int32)
    cid = np.where(s_chr < 5, s_chr, -1).astype(np.int32)
    db_idx, s_idx = db.intersect(cid, s_pos, algo)
    names = np.array(["1", "2", "3", "4", "5"])
    labels = orc.db_chromosome_labels(names, regions)
    o1, o2 = orc.get_common_positions(labels, pos, np.array(["Chr%d" % (c + 1) for c in s_chr]), s_pos)
    assert np.array_equal(db_idx, o1) and np.array_equal(s_idx, o2)
    db.close()


def test_join_rejects_unsorted_markers(lib):
    pos, regions = synth.panel_positions(6000)
    db = lib.Database(pos, regions, 3)
    with pytest.raises(lib.SnpmError):
        db.intersect(np.zeros(3, np.int32), np.array([50, 40, 60], np.int32))
    db.close()


# ---- A3 + A4: inbred ----------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,wei_key,skip", [("pl", "wei", False), ("pl_skip", "wei", True), ("hard", "wei_hard", False)])
def test_inbred_workflow_files_and_arrays(lib, small_geno, sample_inbred, golden_outputs, tmp_path, tag, wei_key, skip):
    from snpmatch_b200.core import parsers, snpmatch
    g = load_golden("inbred_%s.npz" % tag)
    s = sample_inbred
    inp = parsers.ParseInputs("")
    inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s[wei_key], s["dp"])
    out = str(tmp_path / ("inbred_" + tag))
    gt = snpmatch.Genotyper(inp, small_geno, out, run_genotyper=True, skip_db_hets=skip)
    r = gt.result
    assert np.array_equal(gt.commonSNPs[0], g["common_db"]) and np.array_equal(gt.commonSNPs[1], g["common_s"])
    assert r.num_snps == int(g["num_snps"]) and r.overlap == float(g["overlap"])
    assert np.array_equal(r.scores, g["scores"]) and np.array_equal(r.ninfo, g["ninfo"])
    np.testing.assert_allclose(r.probabilies, g["probs"], rtol=0, atol=0, equal_nan=True)
    np.testing.assert_allclose(r.likelis, g["likelis"], rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(r.lrts, g["lrts"], rtol=RTOL, equal_nan=True)
    want = golden_outputs["inbred_" + tag]
    _compare_tables(open(out + ".scores.txt").read(), want["scores.txt"], int_cols=(1, 2, 6), float_cols=(3, 4, 5, 7))
    _json_close(json.loads(open(out + ".matches.json").read()), json.loads(want["matches.json"]))


def _json_close(a, b, rtol=RTOL):
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), (sorted(a), sorted(b))
        for k in a:
            _json_close(a[k], b[k], rtol)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _json_close(x, y, rtol)
    elif isinstance(a, float) or isinstance(b, float):
        if a is None or b is None:
            assert a is b
        elif np.isnan(a) or np.isnan(b):
            assert np.isnan(a) and np.isnan(b)
        else:
            assert abs(a - b) <= rtol * abs(b), (a, b)
    else:
        assert a == b, (a, b)
    return True


def _compare_tables(got, want, int_cols, float_cols, header=False):
    g_lines, w_lines = got.strip("\n").split("\n"), want.strip("\n").split("\n")
    assert len(g_lines) == len(w_lines)
    if header:
        assert g_lines[0] == w_lines[0]
        g_lines, w_lines = g_lines[1:], w_lines[1:]
    for gl, wl in zip(g_lines, w_lines):
        gf, wf = gl.split("\t"), wl.split("\t")
        assert len(gf) == len(wf) and gf[0] == wf[0]
        for c in int_cols:
            assert gf[c] == wf[c], (gl, wl)
        for c in float_cols:
            if wf[c] in ("", "nan"):
                assert gf[c] == wf[c], (gl, wl)
            else:
                assert abs(float(gf[c]) - float(wf[c])) <= RTOL * abs(float(wf[c])), (gl, wl)


# ---- A5 + A6 + A7: cross ------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,sample_name,wei_key,skip", [("pl", "cross", "wei", False), ("hard_skip", "cross", "wei_hard", True),
                                                          ("inbredlike", "inbred", "wei", False)])
def test_cross_workflow_files_and_arrays(lib, small_geno, sample_inbred, sample_cross, golden_outputs, tmp_path, tag, sample_name,
                                         wei_key, skip):
    from snpmatch_b200.core import csmatch, parsers
    g = load_golden("cross_%s.npz" % tag)
    s = sample_cross if sample_name == "cross" else sample_inbred
    inp = parsers.ParseInputs("")
    inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s[wei_key], s["dp"])
    out = str(tmp_path / ("cross_" + tag))
    ci = csmatch.CrossIdentifier(inp, small_geno, "athaliana_tair10", 300000, out, run_identifier=True, skip_db_hets=skip)
    r = ci.result
    n_acc = 40
    assert r.num_snps == int(g["num_snps"]) and r.overlap == float(g["overlap"])
    assert np.array_equal(r.matchedTarInd, g["matchedTarInd"])
    assert np.array_equal(r.winds_chrs, g["winds_chrs"])
    assert np.array_equal(r.accs, g["accs"])
    assert np.array_equal(r.ninfo, g["ninfo"])
    assert np.array_equal(r.scores[:n_acc], g["scores"][:n_acc])                       # truncated window totals: exact
    np.testing.assert_allclose(r.scores[n_acc:], g["scores"][n_acc:], rtol=RTOL)        # simulated F1 rows: float sums
    np.testing.assert_allclose(r.likelis, g["likelis"], rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(r.lrts, g["lrts"], rtol=RTOL, equal_nan=True)
    want = golden_outputs["cross_" + tag]
    _compare_tables(open(out + ".windowscore.txt").read(), want["windowscore.txt"], int_cols=(1, 2, 6, 7), float_cols=(3, 4, 5), header=True)
    _compare_tables(open(out + ".scores.txt").read(), want["scores.txt"], int_cols=(2, 6), float_cols=(1, 3, 4, 5, 7))
    _json_close(json.loads(open(out + ".scores.txt.matches.json").read()), json.loads(want["scores.txt.matches.json"]))
    assert os.path.exists(out + ".matches.json") == ("matches.json" in want)
    if "matches.json" in want:
        _json_close(json.loads(open(out + ".matches.json").read()), json.loads(want["matches.json"]))


def test_window_scores_bit_exact_vs_oracle(lib, small_geno, small_panel, sample_cross):
    from snpmatch_b200.core import genomes, snpmatch
    p, s = small_panel, sample_cross
    w = orc.window_genotyper(p["snps"], p["chrs"], p["chr_regions"], p["positions"], s["chrs"], s["pos"], s["wei"],
                             synth.TAIR10_CHRS, synth.TAIR10_CHRLEN, 300000, False)
    gen = genomes.Genome("athaliana_tair10")
    cnt, off, n_w, _ = gen.window_layout(p["chrs"], 300000)
    order, cid, pos = small_geno.prepare_markers(s["chrs"], s["pos"], style="genome")
    b = lib.Batch(small_geno.db, [0, len(pos)], cid, pos, s["wei"][order])
    b.run_windows(False, 300000, cnt, off, n_w, snpmatch.identity_kmax_table(500, 0.02))
    b.epilogue()
    tot = b.fetch()
    win = b.fetch_windows()
    assert n_w == w.n_windows == 399
    assert np.array_equal(tot["score"][0], w.tot_score) and np.array_equal(tot["ninfo"][0], w.tot_ninfo)
    assert int(tot["m"][0]) == w.num_snps
    seen = np.zeros(n_w, bool)
    for widx, sc, ni in w.windows:
        assert np.array_equal(win["score"][widx - 1], sc), "window %d" % widx
        assert np.array_equal(win["ninfo"][widx - 1], ni)
        lik, lr, ident, num_amb, keep = orc.window_epilogue(sc, ni, 0.02)
        np.testing.assert_allclose(win["L"][widx - 1], lik, rtol=RTOL, equal_nan=True)
        np.testing.assert_allclose(win["LR"][widx - 1], lr, rtol=RTOL, equal_nan=True)
        assert np.array_equal(win["identical"][widx - 1], ident)
        assert int(win["num_amb"][widx - 1]) == num_amb
        seen[widx - 1] = True
    assert np.array_equal(win["nrows"] > 0, seen)
    # the surviving rows, compacted on the device (csmatch.py:57-60): exactly the cells the oracle keeps, in accession order
    rows = b.fetch_window_rows()
    n_acc = p["snps"].shape[1]
    assert np.array_equal(rows["num_amb"], win["num_amb"]) and np.array_equal(rows["nrows"], win["nrows"])
    assert np.array_equal(rows["matched_s_idx"], win["matched_s_idx"])
    expect_off = [0]
    for widx in range(1, n_w + 1):
        amb = int(win["num_amb"][widx - 1])
        k = np.flatnonzero(win["LR"][widx - 1] < snpmatch.lr_thres) if (win["nrows"][widx - 1] > 0 and 1 <= amb < n_acc) else np.zeros(0, int)
        lo, hi = rows["row_off"][widx - 1], rows["row_off"][widx]
        assert hi - lo == len(k), "window %d" % widx
        assert np.array_equal(rows["acc"][lo:hi], k)
        assert np.array_equal(rows["score"][lo:hi], win["score"][widx - 1][k])
        assert np.array_equal(rows["ninfo"][lo:hi], win["ninfo"][widx - 1][k])
        assert np.array_equal(rows["L"][lo:hi], win["L"][widx - 1][k])
        assert np.array_equal(rows["identical"][lo:hi], win["identical"][widx - 1][k])
        expect_off.append(expect_off[-1] + len(k))
    assert rows["row_off"].tolist() == expect_off and expect_off[-1] > 0
    b.close()


def test_vcf701_repo_config(lib, golden_outputs, tmp_path):
    """BASELINE configs[0] stand-in: the reference's sample VCF (parsed weights committed as a fixture)
    against a synthetic database over its positions; inbred and cross outputs vs the reference's."""
    from snpmatch_b200.core import csmatch, parsers, snp_genotype, snpmatch
    p = load_golden("vcf701_panel.npz")
    s = load_golden("vcf701_sample.npz")
    g = snp_genotype.Genotype.from_arrays(p["snps"], p["positions"], p["chrs"], p["chr_regions"], p["accessions"])
    inp = parsers.ParseInputs("")
    inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s["wei"], s["dp"])
    out = str(tmp_path / "vcf701")
    snpmatch.Genotyper(inp, g, out)
    want = golden_outputs["vcf701_inbred"]
    _compare_tables(open(out + ".scores.txt").read(), want["scores.txt"], int_cols=(1, 2, 6), float_cols=(3, 4, 5, 7))
    _json_close(json.loads(open(out + ".matches.json").read()), json.loads(want["matches.json"]))
    csmatch.CrossIdentifier(inp, g, "athaliana_tair10", 300000, out + "_cross")
    want = golden_outputs["vcf701_cross"]
    _compare_tables(open(out + "_cross.windowscore.txt").read(), want["windowscore.txt"], int_cols=(1, 2, 6, 7), float_cols=(3, 4, 5), header=True)
    _compare_tables(open(out + "_cross.scores.txt").read(), want["scores.txt"], int_cols=(2, 6), float_cols=(1, 3, 4, 5, 7))
    assert not os.path.exists(out + "_cross.matches.json")
    g.close()


# ---- batches ------------------------------------------------------------------------------------------
def test_batch_of_samples_equals_single_runs(lib):
    n_rows, n_acc = 60000, 1135
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=3 + 11 * i, n_db=nd, n_extra=ne, seed=900 + i)
               for i, (nd, ne) in enumerate([(2500, 200), (0, 50), (999, 0), (1000, 1), (1001, 7), (4321, 300)])]
    offs = np.concatenate([[0], np.cumsum([len(s["pos"]) for s in samples])])
    b = lib.Batch(db, offs, np.concatenate([s["chr_ix"] for s in samples]), np.concatenate([s["pos"] for s in samples]),
                  np.concatenate([s["wei"] for s in samples]))
    b.run()
    b.epilogue()
    r = b.fetch()
    for i, s in enumerate(samples):
        rows = np.sort(s["rows"]) if "rows" in s else None
        db_idx, s_idx = b.fetch_pairs(i)
        codes = synth.panel_codes(synth.SEED_PANEL, db_idx, n_acc)
        score = np.zeros(n_acc)
        ninfo = np.zeros(n_acc, dtype=np.int64)
        for j in range(0, len(db_idx), 1000):
            t_s, t_n = orc.match_gts_accs(s["wei"][s_idx[j:j + 1000]], codes[j:j + 1000])
            score = score + t_s
            ninfo = ninfo + t_n
        assert int(r["m"][i]) == len(db_idx)
        assert np.array_equal(pos[db_idx].astype(np.int64), s["pos"][s_idx])
        assert np.array_equal(r["score"][i], score) and np.array_equal(r["ninfo"][i], ninfo)
        assert np.array_equal(r["matches"][i], score.astype(np.int64))
        lik, lr = orc.calculate_likelihoods(score.astype(np.int64), ninfo)
        np.testing.assert_allclose(r["L"][i], lik, rtol=RTOL, equal_nan=True)
        np.testing.assert_allclose(r["LR"][i], lr, rtol=RTOL, equal_nan=True)
        if len(db_idx) > 500:
            assert int(np.nanargmin(r["L"][i])) == 3 + 11 * i
    # dictionary-coded weights expand to the same bits
    idx, table = lib.index_weights(np.concatenate([s["wei"] for s in samples]))
    assert len(table) < 2000
    b.upload_indexed(offs, np.concatenate([s["chr_ix"] for s in samples]), np.concatenate([s["pos"] for s in samples]), idx, table)
    b.run()
    b.epilogue()
    r_idx = b.fetch()
    for k in ("score", "matches", "ninfo", "m", "L", "LR", "prob"):
        assert np.array_equal(r_idx[k], r[k], equal_nan=True), k
    # re-upload into the same batch object
    s = samples[0]
    b.upload([0, len(s["pos"])], s["chr_ix"], s["pos"], s["wei_hard"])
    b.run(skip_db_hets=True)
    b.epilogue()
    r1 = b.fetch()
    one = lib.Batch(db, [0, len(s["pos"])], s["chr_ix"], s["pos"], s["wei_hard"])
    one.run(skip_db_hets=True)
    one.epilogue()
    r2 = one.fetch()
    for k in ("score", "matches", "ninfo", "m"):
        assert np.array_equal(r1[k], r2[k])
    t = one.timings()
    assert t["launches"] >= 6 and t["total_ms"] > 0
    one.close(); b.close(); db.close()


@pytest.mark.parametrize("skip", [False, True])
def test_popcount_kernel_equals_fp64_kernel_for_called_genotypes(lib, skip):
    n_rows, n_acc = 80000, 1135
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    samples = [synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=5 + 7 * i, n_db=nd, n_extra=ne, seed=700 + i, het=0.05)
               for i, (nd, ne) in enumerate([(3333, 100), (1, 0), (1000, 10), (2049, 5), (0, 3)])]
    offs = np.concatenate([[0], np.cumsum([len(s["pos"]) for s in samples])])
    wei = np.concatenate([s["wei_hard"] for s in samples])
    assert lib.weights_are_one_hot(wei) and not lib.weights_are_one_hot(samples[0]["wei"])
    b = lib.Batch(db, offs, np.concatenate([s["chr_ix"] for s in samples]), np.concatenate([s["pos"] for s in samples]), wei)
    res = {}
    for mode in (lib.KERNEL_FP64, lib.KERNEL_POPCOUNT):
        b.run(skip_db_hets=skip, kernel_mode=mode)
        b.epilogue()
        res[mode] = {k: v.copy() for k, v in b.fetch().items()}
    for k in ("score", "matches", "ninfo", "m", "prob", "L", "LR"):
        assert np.array_equal(res[0][k], res[1][k], equal_nan=True), k
    # and against the oracle for one sample
    s = samples[0]
    db_idx, s_idx = b.fetch_pairs(0)
    ref_s, ref_n = np.zeros(n_acc), np.zeros(n_acc, dtype=np.int64)
    codes = synth.panel_codes(synth.SEED_PANEL, db_idx, n_acc)
    for j in range(0, len(db_idx), 1000):
        t_s, t_n = orc.match_gts_accs(s["wei_hard"][s_idx[j:j + 1000]], codes[j:j + 1000], skip)
        ref_s, ref_n = ref_s + t_s, ref_n + t_n
    assert np.array_equal(res[1]["score"][0], ref_s) and np.array_equal(res[1]["ninfo"][0], ref_n)
    b.close()
    # likelihood weights are refused by the popcount kernel
    s = samples[0]
    b = lib.Batch(db, [0, len(s["pos"])], s["chr_ix"], s["pos"], s["wei"])
    b.run(kernel_mode=lib.KERNEL_POPCOUNT)
    with pytest.raises(lib.SnpmError):
        b.wait()
    b.close()
    db.close()


def test_row_filter_refine_path(lib, small_geno, small_panel, sample_inbred):
    p, s = small_panel, sample_inbred
    keep_rows = np.arange(0, 6000, 3)
    ref = orc.genotyper(p["snps"], p["chrs"], p["chr_regions"], p["positions"], s["chrs"], s["pos"], s["wei"], filter_pos_ix=keep_rows)
    order, cid, pos = small_geno.prepare_markers(s["chrs"], s["pos"])
    b = lib.Batch(small_geno.db, [0, len(pos)], cid, pos, s["wei"][order])
    b.set_row_filter(keep_rows)
    b.run()
    b.epilogue()
    r = b.fetch()
    assert int(r["m"][0]) == ref.num_snps
    assert np.array_equal(r["score"][0], ref.score_f64) and np.array_equal(r["ninfo"][0], ref.ninfo)
    b.close()


def test_dense_sample_properties(lib):
    """Full-size shape check by properties: a sample that carries EVERY database row of one accession
    must match that accession perfectly; ninfo equals the accession's called rows."""
    n_rows, n_acc = 300000, 1135
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    rows = np.arange(n_rows)
    code = synth.panel_codes_cols(synth.SEED_PANEL, rows, [42])[:, 0]
    chrom = np.searchsorted(regions[:, 1], rows, side="right").astype(np.int32)
    wei = synth.hard_weights(np.where(code < 0, 0, code))
    for algo in (1, 2):
        b = lib.Batch(db, [0, n_rows], chrom, pos, wei)
        b.run(join_algo=algo)
        b.epilogue()
        r = b.fetch()
        called = int((code >= 0).sum())
        assert int(r["m"][0]) == n_rows
        assert int(r["ninfo"][0, 42]) == called and int(r["matches"][0, 42]) == called
        assert r["L"][0, 42] == 1.0 and int(np.nanargmin(r["L"][0])) == 42
        assert r["ninfo"][0].max() <= n_rows and (r["matches"][0] <= r["ninfo"][0]).all()
        b.close()
    db.close()


def test_wide_panel_20k_accessions(lib):
    """BASELINE configs[4] shape at reduced row count: 20 000 accessions (626-word rows, 20 word slices per segment);
    fp64 and popcount kernels against the oracle on the rows they touch."""
    n_rows, n_acc = 40000, 20000
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    s = synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=12345, n_db=2100, n_extra=100, seed=42)
    b = lib.Batch(db, [0, len(s["pos"])], s["chr_ix"], s["pos"], s["wei"])
    b.run()
    b.epilogue()
    r = b.fetch()
    db_idx, s_idx = b.fetch_pairs(0)
    codes = synth.panel_codes(synth.SEED_PANEL, db_idx, n_acc)
    score, ninfo = np.zeros(n_acc), np.zeros(n_acc, dtype=np.int64)
    for j in range(0, len(db_idx), 1000):
        t_s, t_n = orc.match_gts_accs(s["wei"][s_idx[j:j + 1000]], codes[j:j + 1000])
        score, ninfo = score + t_s, ninfo + t_n
    assert np.array_equal(r["score"][0], score) and np.array_equal(r["ninfo"][0], ninfo)
    assert int(np.nanargmin(r["L"][0])) == 12345
    b.upload([0, len(s["pos"])], s["chr_ix"], s["pos"], s["wei_hard"])
    b.run(kernel_mode=lib.KERNEL_POPCOUNT)
    b.epilogue()
    r = b.fetch()
    score, ninfo = np.zeros(n_acc), np.zeros(n_acc, dtype=np.int64)
    for j in range(0, len(db_idx), 1000):
        t_s, t_n = orc.match_gts_accs(s["wei_hard"][s_idx[j:j + 1000]], codes[j:j + 1000])
        score, ninfo = score + t_s, ninfo + t_n
    assert np.array_equal(r["score"][0], score) and np.array_equal(r["ninfo"][0], ninfo)
    b.close()
    db.close()


def test_cross_full_panel_properties(lib):
    """BASELINE configs[2] shape by properties: windows of a sample drawn from ONE accession; every non-empty
    window's best likelihood belongs to that accession, window counts add up to the totals, F1 pairs with the true
    accession as a parent carry its called sites."""
    from snpmatch_b200.core import genomes, snpmatch
    n_rows, n_acc = 400000, 1135
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    s = synth.make_sample(pos, regions, synth.TAIR10_CHRS, n_acc, true_acc=77, n_db=20000, n_extra=500, seed=43, err=0.0, het=0.0)
    gen = genomes.Genome("athaliana_tair10")
    cnt, off, n_w, _ = gen.window_layout(np.array(synth.TAIR10_CHRS), 300000)
    b = lib.Batch(db, [0, len(s["pos"])], s["chr_ix"], s["pos"], s["wei_hard"])
    b.run_windows(False, 300000, cnt, off, n_w, snpmatch.identity_kmax_table(2000, 0.02))
    b.epilogue()
    tot = b.fetch()
    w = b.fetch_windows()
    assert n_w == 399 and int(w["nrows"].sum()) == int(tot["m"][0]) == len(w["matched_s_idx"])
    assert np.array_equal(w["ninfo"].astype(np.int64).sum(axis=0), tot["ninfo"][0])
    assert np.array_equal(w["score"].sum(axis=0), tot["score"][0])            # 0/1 weights: exact in any order
    live = np.flatnonzero(w["nrows"] > 0)
    assert len(live) > 380
    assert np.all(w["score"][live, 77] == w["ninfo"][live, 77])               # perfect match in every window
    assert np.all(w["L"][live, 77] == 1.0) and np.all(w["identical"][live, 77] == 1)
    assert int(np.nanargmin(tot["L"][0])) == 77
    top = np.argsort(-tot["prob"][0])[:10]
    f_score, f_ninfo = b.f1_pairs(top)
    assert len(f_score) == 45 and np.all(f_ninfo <= int(tot["m"][0])) and np.all(f_score <= f_ninfo)
    b.close()
    db.close()


def test_segregating_rows_and_refine(lib, small_geno, small_panel, sample_inbred, tmp_path):
    from snpmatch_b200.core import parsers, snpmatch
    p = small_panel
    for sel in (np.array([7, 9]), np.array([0, 3, 11, 17, 21, 30, 39]), np.array([11, 12])):
        got = snpmatch.identify_segregating_snps(small_geno, sel)
        assert np.array_equal(got, orc.segregating_rows(p["snps"], sel))
    assert snpmatch.identify_segregating_snps(small_geno, np.arange(25)) is None
    # --refine end to end: accession 9 is a near-copy of accession 7 in the golden panel
    s = sample_inbred
    inp = parsers.ParseInputs("")
    inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s["wei"], s["dp"])
    out = str(tmp_path / "refine")
    gt = snpmatch.Genotyper(inp, small_geno, out, run_genotyper=False)
    gt.filter_tophits()
    assert os.path.exists(out + ".scores.txt")
    with np.errstate(invalid="ignore"):
        top = np.flatnonzero(gt.result.lrts < snpmatch.lr_thres)
    if 1 < len(top) <= 20:
        seg = orc.segregating_rows(p["snps"], top)
        ref = orc.genotyper(p["snps"], p["chrs"], p["chr_regions"], p["positions"], s["chrs"], s["pos"], s["wei"], filter_pos_ix=seg)
        with np.errstate(invalid="ignore"):
            keep = np.setdiff1d(np.arange(40), np.flatnonzero(gt.result.lrts >= snpmatch.lr_thres))   # nan ratios stay, as in the reference
        assert os.path.exists(out + ".refined.scores.txt")
        assert np.array_equal(gt.result_fine.scores, ref.scores[keep]) and np.array_equal(gt.result_fine.ninfo, ref.ninfo[keep])
        assert gt.result_fine.num_snps == ref.num_snps


@pytest.mark.parametrize("n_acc,S,K,skip", [(40, 70, 300, False), (1135, 130, 1000, False), (300, 64, 33, True), (129, 5, 1, False)])
def test_shared_panel_tensor_core_batch(lib, n_acc, S, K, skip):
    """A9: one-hot int8 GEMM on tcgen05 vs the oracle's matchGTsAccs per sample (integers, exact)."""
    n_rows = 20000
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    rng = np.random.default_rng(S * 1000 + K)
    rows = np.sort(rng.choice(n_rows, size=K, replace=False))
    codes = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(S, K), p=[0.55, 0.3, 0.05, 0.1])
    r = db.score_shared_panel(rows, codes, skip_db_hets=skip)
    panel = synth.panel_codes(synth.SEED_PANEL, rows, n_acc)
    for s in range(S):
        have = np.flatnonzero(codes[s] < 3)
        wei = synth.hard_weights(codes[s][have].astype(np.int8))
        sc, ni = orc.match_gts_accs(wei, panel[have], skip)
        assert np.array_equal(r["matches"][s], sc.astype(np.int64)), "sample %d" % s
        assert np.array_equal(r["ninfo"][s], ni), "sample %d" % s
        lik, lr = orc.calculate_likelihoods(sc.astype(np.int64), ni)
        np.testing.assert_allclose(r["L"][s], lik, rtol=RTOL, equal_nan=True)
        np.testing.assert_allclose(r["LR"][s], lr, rtol=RTOL, equal_nan=True)
    assert r["gemm_ms"] > 0
    db.close()


@pytest.mark.parametrize("n_acc,K,skip", [(1135, 1001, False), (300, 34, True), (129, 3, False), (257, 64, False)])
def test_shared_panel_object_packed_and_plain(lib, n_acc, K, skip):
    """A9 as an object (snpm_panel_*): the panel operand is expanded once and re-used by batches of different sizes; 2-bit
    packed codes (K not a multiple of 4 included, garbage in the spare bits) and uint8 codes give the oracle's integers."""
    n_rows = 20000
    pos, regions = synth.panel_positions(n_rows)
    db = lib.Database(pos, regions, n_acc)
    db.fill_synthetic(synth.SEED_PANEL)
    rng = np.random.default_rng(7000 + K)
    rows = np.sort(rng.choice(n_rows, size=K, replace=False))
    panel = synth.panel_codes(synth.SEED_PANEL, rows, n_acc)
    sp = db.shared_panel(rows, skip_db_hets=skip)
    for S in (70, 5, 130):                                          # grows, shrinks, grows: scratch is re-used
        codes = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(S, K), p=[0.55, 0.3, 0.05, 0.1])
        packed = lib.pack_codes2(codes)
        assert packed.shape == (S, (K + 3) // 4)
        if K % 4:
            packed[:, -1] &= np.uint8((1 << (2 * (K % 4))) - 1)     # spare bits 0 = would read as ref calls if not masked
        r_plain = sp.score(codes, likelihoods=True)
        r_packed = sp.score(packed, packed=True, likelihoods=True)
        for s in range(S):
            have = np.flatnonzero(codes[s] < 3)
            sc, ni = orc.match_gts_accs(synth.hard_weights(codes[s][have].astype(np.int8)), panel[have], skip)
            for r in (r_plain, r_packed):
                assert r["matches"].dtype == np.int32
                assert np.array_equal(r["matches"][s], sc.astype(np.int64)) and np.array_equal(r["ninfo"][s], ni), "sample %d" % s
            lik, lr = orc.calculate_likelihoods(sc.astype(np.int64), ni)
            np.testing.assert_allclose(r_packed["L"][s], lik, rtol=RTOL, equal_nan=True)
            np.testing.assert_allclose(r_packed["LR"][s], lr, rtol=RTOL, equal_nan=True)
        assert np.array_equal(r_plain["L"], r_packed["L"], equal_nan=True)
        r_int = sp.score(packed, packed=True)                       # no likelihoods: integers only
        assert np.array_equal(r_int["matches"], r_packed["matches"]) and "L" not in r_int
        assert r_int["gemm_ms"] > 0 and r_int["device_ms"] >= r_int["gemm_ms"]
    one = db.score_shared_panel(rows, codes, skip_db_hets=skip)      # the one-shot form (int64) agrees
    assert np.array_equal(one["matches"], r_packed["matches"]) and np.array_equal(one["ninfo"], r_packed["ninfo"])
    with pytest.raises(Exception):
        lib.SharedPanel(db, np.array([n_rows + 5], dtype=np.int64))
    sp.close()
    db.close()


def test_genotype_batch_equals_per_sample_genotyper(lib, small_geno, small_panel, tmp_path):
    """Batched tensor-core mode through the Python mirror vs Genotyper run sample by sample (called genotypes)."""
    from snpmatch_b200.core import batch, parsers, snpmatch
    p = small_panel
    inputs = []
    for i in range(9):
        s = synth.make_sample(p["positions"], p["chr_regions"], p["chrs"], 40, true_acc=3 + 4 * i, n_db=800 + 50 * i, n_extra=40, seed=300 + i)
        inp = parsers.ParseInputs("")
        inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s["wei_hard"], s["dp"])
        inputs.append(inp)
    for skip in (False, True):
        results, info = batch.genotype_batch(small_geno, inputs, skip_db_hets=skip)
        assert info["panel_markers"] > 800 and info["gemm_ms"] > 0
        for i, inp in enumerate(inputs):
            one = snpmatch.Genotyper(inp, small_geno, str(tmp_path / ("b%d" % i)), run_genotyper=False, skip_db_hets=skip).genotyper()
            got = results[i]
            assert np.array_equal(got.scores, one.scores) and np.array_equal(got.ninfo, one.ninfo)
            assert got.num_snps == one.num_snps and got.overlap == one.overlap
            got.get_likelihoods(); one.get_likelihoods()
            np.testing.assert_allclose(got.likelis, one.likelis, rtol=RTOL, equal_nan=True)
            np.testing.assert_allclose(got.lrts, one.lrts, rtol=RTOL, equal_nan=True)
            if not skip and 3 + 4 * i not in (9, 11):      # 11 is all-missing and 9 a near-copy of 7 in the golden panel
                assert int(np.nanargmin(got.likelis)) == 3 + 4 * i


def test_command_line_inbred_and_cross(lib, small_geno, sample_inbred, golden_outputs, tmp_path):
    """The `snpmatch inbred` / `cross` command line on a packed database file and a BED sample."""
    import snpmatch_b200
    db_path = str(tmp_path / "panel.npz")
    small_geno.save_packed(db_path)
    s = sample_inbred
    bed = tmp_path / "sample.bed"
    with open(bed, "w") as fh:
        for c, pos, gt in zip(s["chrs"], s["pos"], s["gt"]):
            fh.write("%s\t%d\t%s\n" % (c, pos, gt))
    out = str(tmp_path / "cli")
    assert snpmatch_b200.main(["inbred", "-i", str(bed), "-d", db_path, "-o", out]) == 0
    want = golden_outputs["inbred_hard"]["scores.txt"].strip("\n").split("\n")
    got = open(out + ".scores.txt").read().strip("\n").split("\n")
    assert len(got) == len(want)
    for g_line, w_line in zip(got, want):
        gf, wf = g_line.split("\t"), w_line.split("\t")
        assert gf[:3] == wf[:3] and gf[6] == wf[6]                 # accession, matches, ninfo, num_snps
        assert gf[7] == ""                                          # BED carries no depth: nan instead of the reference's crash (A.8 Q5)
    js = json.loads(open(out + ".matches.json").read())
    assert js["interpretation"]["case"] == json.loads(golden_outputs["inbred_hard"]["matches.json"])["interpretation"]["case"]
    assert snpmatch_b200.main(["cross", "-i", str(bed), "-d", db_path, "-o", out + "_x", "-b", "300000"]) == 0
    assert os.path.exists(out + "_x.windowscore.txt") and os.path.exists(out + "_x.scores.txt")
    with pytest.raises(SystemExit):                                 # die(): missing input file -> exit 1, as the reference
        snpmatch_b200.main(["inbred", "-i", str(tmp_path / "missing.bed"), "-d", db_path])


def test_empty_row_filter_keeps_nothing(lib, small_geno, small_panel, sample_inbred, tmp_path):
    """Genotyper.genotyper(filter_pos_ix=<empty>) (snpmatch.py:211-216): the reference is left without any common SNP, so
    every score and count is 0 (not the unfiltered scores); None clears the filter again."""
    from snpmatch_b200.core import parsers, snpmatch
    p, s = small_panel, sample_inbred
    ref = orc.genotyper(p["snps"], p["chrs"], p["chr_regions"], p["positions"], s["chrs"], s["pos"], s["wei"],
                        filter_pos_ix=np.zeros(0, dtype=np.int64))
    assert ref.num_snps == 0
    order, cid, pos = small_geno.prepare_markers(s["chrs"], s["pos"])
    b = lib.Batch(small_geno.db, [0, len(pos)], cid, pos, s["wei"][order])
    b.run()
    b.epilogue()
    full = {k: v.copy() for k, v in b.fetch().items()}
    b.set_row_filter(np.zeros(0, dtype=np.int64))
    b.run()
    b.epilogue()
    r = b.fetch()
    assert int(r["m"][0]) == 0 and np.all(r["score"][0] == 0.0) and np.all(r["ninfo"][0] == 0) and np.all(np.isnan(r["L"][0]))
    b.set_row_filter(None)
    b.run()
    b.epilogue()
    again = b.fetch()
    assert int(again["m"][0]) == int(full["m"][0]) and np.array_equal(again["score"][0], full["score"][0])
    b.close()
    inp = parsers.ParseInputs("")
    inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s["wei"], s["dp"])
    gt = snpmatch.Genotyper(inp, small_geno, str(tmp_path / "e"), run_genotyper=False)
    res = gt.genotyper(filter_pos_ix=np.zeros(0, dtype=np.int64))
    assert res.num_snps == 0 and np.all(res.scores == 0)


@pytest.mark.parametrize("chunk", [1, 37, 500, 1000, 2500, 100000])
def test_genotyper_chunk_size_is_part_of_the_summation_order(lib, small_geno, small_panel, sample_inbred, tmp_path, chunk):
    """Genotyper(chunk_size=...) (snpmatch.py:173,218): the reference adds chunk sums, so its fp64 score depends on the chunk
    size in the last bits; the order-exact kernel follows any chunk size bit for bit."""
    from snpmatch_b200.core import parsers, snpmatch
    p, s = small_panel, sample_inbred
    ref = orc.genotyper(p["snps"], p["chrs"], p["chr_regions"], p["positions"], s["chrs"], s["pos"], s["wei"], chunk_size=chunk)
    inp = parsers.ParseInputs("")
    inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s["wei"], s["dp"])
    res = snpmatch.Genotyper(inp, small_geno, str(tmp_path / "c"), run_genotyper=False, chunk_size=chunk).genotyper()
    assert res.num_snps == ref.num_snps and np.array_equal(res.ninfo, ref.ninfo)
    assert np.array_equal(np.asarray(res.scores), ref.scores)
    b = lib.Batch(small_geno.db, [0, 0], np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3)))
    order, cid, pos = small_geno.prepare_markers(s["chrs"], s["pos"])
    b.set_chunk_rows(chunk)
    b.upload([0, len(pos)], cid, pos, s["wei"][order])
    b.run()
    b.epilogue()
    assert np.array_equal(b.fetch()["score"][0], ref.score_f64), "fp64 scores must follow the reference's chunked sums"
    b.close()
    # back to the default for the database's shared scratch batch
    snpmatch.Genotyper(inp, small_geno, str(tmp_path / "d"), run_genotyper=False).genotyper()


def test_int8_codes_outside_the_reference_range(lib):
    """Negative codes are missing calls (the reference masks every value < 0); codes above 2 do not exist in its databases
    (makedb.py:59) and are refused instead of being folded into 0..3."""
    n_rows, n_acc = 40, 70
    pos = (np.arange(n_rows, dtype=np.int32) + 1) * 10
    regions = np.array([[0, n_rows]], dtype=np.int64)
    rng = np.random.default_rng(3)
    snps = rng.integers(-1, 3, size=(n_rows, n_acc)).astype(np.int8)
    weird = snps.copy()
    weird[snps == -1] = rng.choice(np.array([-2, -7, -128], dtype=np.int8), size=int((snps == -1).sum()))
    db = lib.Database(pos, regions, n_acc)
    db.load_int8(weird)
    assert np.array_equal(db.read_rows(np.arange(n_rows)), snps)
    bad = snps.copy()
    bad[5, 9] = 3
    with pytest.raises(lib.SnpmError):
        db.load_int8(bad)
    with pytest.raises(lib.SnpmError):
        lib.match_gts_accs(np.full((n_rows, 3), 0.5), bad)
    s, n = lib.match_gts_accs(np.full((n_rows, 3), 0.5), weird)
    rs, rn = orc.match_gts_accs(np.full((n_rows, 3), 0.5), weird)
    assert np.array_equal(s, rs) and np.array_equal(n, rn)
    db.close()


def test_cross_on_two_row_shards_equals_one_database(lib, small_panel, sample_cross):
    """SURVEY 8(e) for `cross`: the panel cut into two SNP-row shards (both held on this one GPU, the exchange step done by
    hand), per-window partials summed through the packed buffer of snpm_batch_run_windows_begin, then the second half on
    both — against the unsharded run: integers and identity calls ==, float window scores == except in the window that
    straddles the shard boundary (two in-order partial sums instead of one: rtol 1e-12), and against the oracle."""
    import torch
    from snpmatch_b200 import sharding
    from snpmatch_b200.core import genomes, snpmatch
    p, s = small_panel, sample_cross
    n_rows, n_acc = len(p["positions"]), p["snps"].shape[1]
    gen = genomes.Genome("athaliana_tair10")
    cnt, off, n_w, _ = gen.window_layout(p["chrs"], 300000)
    kmax = snpmatch.identity_kmax_table(4000, 0.02)
    first = {str(c): i for i, c in enumerate(p["chrs"])}
    cid = np.array([first[str(c).replace("Chr", "").replace("chr", "")] for c in s["chrs"]], dtype=np.int32)
    pos = s["pos"].astype(np.int32)
    # reference: one database
    db = lib.Database(p["positions"], p["chr_regions"], n_acc)
    db.load_int8(p["snps"])
    b = lib.Batch(db, [0, len(pos)], cid, pos, s["wei"])
    b.run_windows(False, 300000, cnt, off, n_w, kmax)
    b.epilogue()
    tot = {k: v.copy() for k, v in b.fetch().items()}
    ref = b.fetch_windows()
    ref_rows = b.fetch_window_rows()
    f1_ref = b.f1_pairs(np.argsort(-tot["prob"][0])[:10])
    b.close()
    db.close()
    # two shards with a boundary in the middle of a chromosome (and of a window)
    cut = int(p["chr_regions"][1][0] + (p["chr_regions"][1][1] - p["chr_regions"][1][0]) // 3)
    shards = []
    for r0, r1 in ((0, cut), (cut, n_rows)):
        d = lib.Database(p["positions"][r0:r1], sharding.local_regions(p["chr_regions"], r0, r1), n_acc, row0_global=r0)
        d.load_int8(p["snps"][r0:r1])
        i0, i1 = sharding.shard_marker_range(cid, pos, p["chr_regions"], p["positions"], r0, r1)
        bb = lib.Batch(d, [0, i1 - i0], cid[i0:i1], pos[i0:i1], s["wei"][i0:i1])
        shards.append((d, bb, i0))
    views = []
    for d, bb, _ in shards:
        ptr, n = bb.run_windows_begin(False, 300000, cnt, off, n_w, kmax)
        views.append(sharding.device_view(ptr, n, torch.device("cuda", 0)))
    torch.cuda.synchronize()
    total = views[0].clone() + views[1]                      # the all-reduce, by hand
    for v in views:
        v.copy_(total)
    torch.cuda.synchronize()
    matched = []
    for d, bb, i0 in shards:
        bb.run_windows_finish()
        bb.epilogue()
        t2 = bb.fetch()
        w2 = bb.fetch_windows()
        r2 = bb.fetch_window_rows()
        matched.append(w2["matched_s_idx"] + i0)
        assert np.array_equal(w2["ninfo"], ref["ninfo"]) and np.array_equal(w2["nrows"], ref["nrows"])
        assert np.array_equal(w2["identical"], ref["identical"]) and np.array_equal(w2["num_amb"], ref["num_amb"])
        differs = np.flatnonzero((w2["score"] != ref["score"]).any(axis=1))
        assert len(differs) <= 1, "only the window on the shard boundary may differ in the last bits"
        np.testing.assert_allclose(w2["score"], ref["score"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(w2["L"], ref["L"], rtol=RTOL, equal_nan=True)
        assert np.array_equal(t2["ninfo"][0], tot["ninfo"][0]) and np.array_equal(t2["matches"][0], tot["matches"][0])
        assert np.array_equal(r2["row_off"], ref_rows["row_off"]) and np.array_equal(r2["acc"], ref_rows["acc"])
        assert np.array_equal(r2["ninfo"], ref_rows["ninfo"]) and np.array_equal(r2["identical"], ref_rows["identical"])
    assert np.array_equal(np.concatenate(matched), ref["matched_s_idx"])
    # simulated F1s: partial sums of the two shards
    top = np.argsort(-tot["prob"][0])[:10]
    parts = [bb.f1_pairs(top) for _, bb, _ in shards]
    np.testing.assert_allclose(parts[0][0] + parts[1][0], f1_ref[0], rtol=1e-12)
    assert np.array_equal(parts[0][1] + parts[1][1], f1_ref[1])
    # oracle
    o = orc.window_genotyper(p["snps"], p["chrs"], p["chr_regions"], p["positions"], s["chrs"], s["pos"], s["wei"], gen.chrs, gen.chrlen, 300000)
    assert o.num_snps == int(ref["nrows"].sum())
    for widx, sc, ni in o.windows:
        assert np.array_equal(ref["ninfo"][widx - 1], ni) and np.array_equal(ref["score"][widx - 1], sc)
    for d, bb, _ in shards:
        bb.close()
        d.close()
