"""Pins the CPU oracle (oracle/snpmatch_oracle.py) against vectors produced by the
UNMODIFIED reference (oracle/gen_golden.py) and against the reference's own known answers."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import snpmatch_oracle as orc


def test_likelihood_known_answer():
    # tests/test_inbred.py:22 and tests/test_genotype_cross.py:23 of the reference
    assert orc.likeli_test(10, 3) == 122.8361221819443
    assert np.isnan(orc.likeli_test(10, 0))        # tests/test_inbred.py:24
    assert np.isnan(orc.likeli_test(0, 0))
    g = load_golden("epilogue.npz")
    assert float(g["likeli_10_3"]) == 122.8361221819443
    L, _ = orc.calculate_likelihoods(np.array([3]), np.array([10]))
    assert L[0] == 122.8361221819443


def test_match_gts_accs_bit_exact():
    g = load_golden("match_gts_accs.npz")
    for i in range(int(g["n_cases"])):
        snps, wei, skip = g["c%d_snps" % i], g["c%d_wei" % i], bool(g["c%d_skip" % i])
        score, ninfo = orc.match_gts_accs(wei, snps, skip)
        assert np.array_equal(ninfo, g["c%d_ninfo" % i])
        assert np.array_equal(score, g["c%d_score" % i]), "fp64 score not bit-exact in case %d" % i
        if snps.shape[0] <= 300:
            s2, n2 = orc.match_gts_accs_sequential(wei, snps, skip)
            assert np.array_equal(s2, g["c%d_score" % i])
            assert np.array_equal(n2, g["c%d_ninfo" % i])


def test_join_cases():
    g = load_golden("join.npz")
    for i in range(int(g["n_cases"])):
        i1, i2 = orc.get_common_positions(g["j%d_c1" % i], g["j%d_p1" % i], g["j%d_c2" % i], g["j%d_p2" % i])
        assert np.array_equal(i1, g["j%d_i1" % i])
        assert np.array_equal(i2, g["j%d_i2" % i])
    # SURVEY A.1 probed example
    i1, i2 = orc.get_common_positions(g["j0_c1"], g["j0_p1"], g["j0_c2"], g["j0_p2"])
    assert i1.tolist() == [1, 4, 6, 8] and i2.tolist() == [3, 4, 0, 2]


def test_epilogue_sets():
    g = load_golden("epilogue.npz")
    for i in range(int(g["n_sets"])):
        with np.errstate(all="ignore"):
            L, LR = orc.calculate_likelihoods(g["e%d_y" % i], g["e%d_n" % i])
        np.testing.assert_allclose(L, g["e%d_L" % i], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(LR, g["e%d_LR" % i], rtol=1e-12, equal_nan=True)


def test_identity_grid_and_kmax_table():
    g = load_golden("epilogue.npz")
    n, x, xf = g["id_n"], g["id_x"], g["id_xf"]
    assert np.array_equal(orc.test_identity(x, n, error_rate=0.02), g["id_e02"])
    assert np.array_equal(orc.test_identity(xf, n, error_rate=0.02), g["id_e02_f"])
    assert np.array_equal(orc.test_identity(x, n), g["id_default"])
    kmax = orc.identity_kmax_table(2000, 0.02)
    # identical <=> floor(n - x - 1) + 1 <= kmax[n]   (SciPy floors the first argument of sf)
    assert np.array_equal((np.floor(n - x - 1) + 1 <= kmax[n]).astype(int), g["id_e02"])
    assert np.array_equal((np.floor(n - xf - 1) + 1 <= kmax[n]).astype(int), g["id_e02_f"])
    # probed values, SURVEY A.4
    for nn, k in ((1, 0), (2, 0), (3, 1), (5, 1), (10, 1), (20, 2), (50, 3), (100, 5), (200, 7), (500, 15), (1000, 28)):
        assert kmax[nn] == k


@pytest.mark.parametrize("tag,wei_key,skip", [("pl", "wei", False), ("pl_skip", "wei", True), ("hard", "wei_hard", False)])
def test_inbred_workflow(small_panel, sample_inbred, tag, wei_key, skip):
    g = load_golden("inbred_%s.npz" % tag)
    p, s = small_panel, sample_inbred
    r = orc.genotyper(p["snps"], p["chrs"], p["chr_regions"], p["positions"], s["chrs"], s["pos"], s[wei_key], skip)
    assert np.array_equal(r.common[0], g["common_db"]) and np.array_equal(r.common[1], g["common_s"])
    assert r.num_snps == int(g["num_snps"]) and r.overlap == float(g["overlap"])
    assert np.array_equal(r.scores, g["scores"])
    assert np.array_equal(r.ninfo, g["ninfo"])
    np.testing.assert_allclose(r.likelis, g["likelis"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(r.lrts, g["lrts"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(r.probabilities, g["probs"], rtol=0, atol=0, equal_nan=True)


@pytest.mark.parametrize("tag,sample_name,wei_key,skip", [("pl", "cross", "wei", False), ("hard_skip", "cross", "wei_hard", True),
                                                          ("inbredlike", "inbred", "wei", False)])
def test_cross_workflow(small_panel, sample_inbred, sample_cross, golden_outputs, tag, sample_name, wei_key, skip):
    from snpmatch_b200 import synth
    g = load_golden("cross_%s.npz" % tag)
    p = small_panel
    s = sample_cross if sample_name == "cross" else sample_inbred
    w = orc.window_genotyper(p["snps"], p["chrs"], p["chr_regions"], p["positions"], s["chrs"], s["pos"], s[wei_key],
                             synth.TAIR10_CHRS, synth.TAIR10_CHRLEN, 300000, skip)
    n_acc = p["snps"].shape[1]
    assert w.n_windows == 399
    assert w.num_snps == int(g["num_snps"]) and w.overlap == float(g["overlap"])
    assert np.array_equal(w.matched_tar, g["matchedTarInd"])
    assert np.array_equal(w.winds_chrs, g["winds_chrs"])
    # totals: the first A rows of the final table are the truncated window totals
    assert np.array_equal(np.array(w.tot_score, dtype="int").astype(np.float64), g["scores"][:n_acc])
    assert np.array_equal(w.tot_ninfo, g["ninfo"][:n_acc])
    # F1 pass
    probs = orc.probabilities(np.array(w.tot_score, dtype="int"), w.tot_ninfo)
    top = orc.top_hit_accessions(probs)
    labels = orc.db_chromosome_labels(p["chrs"], p["chr_regions"])
    common = orc.get_common_positions(labels, p["positions"], s["chrs"], s["pos"])
    pairs, f_sc, f_ni = orc.f1_pair_scores(p["snps"], common, s[wei_key], top)
    assert np.array_equal(f_sc, g["scores"][n_acc:])
    assert np.array_equal(f_ni, g["ninfo"][n_acc:])
    accs = p["accessions"].astype("U")
    assert [accs[i] + "x" + accs[j] for i, j in pairs] == g["accs"][n_acc:].tolist()
    # window table rows against the reference's windowscore.txt
    lines = golden_outputs["cross_" + tag]["windowscore.txt"].strip("\n").split("\n")[1:]
    got = []
    for (widx, sc, ni) in w.windows:
        lik, lr, ident, num_amb, keep = orc.window_epilogue(sc, ni, 0.02)
        for a in np.flatnonzero(keep):
            got.append((accs[a], int(float(sc[a])), int(ni[a]), sc[a] / ni[a], lik[a], float(ident[a]), num_amb, widx))
    assert len(got) == len(lines)
    for row, line in zip(got, lines):
        f = line.split("\t")
        assert f[0] == row[0] and int(f[1]) == row[1] and int(f[2]) == row[2]
        assert float(f[5]) == row[5] and int(f[6]) == row[6] and int(f[7]) == row[7]
        np.testing.assert_allclose(float(f[3]), row[3], rtol=1e-12)
        np.testing.assert_allclose(float(f[4]), row[4], rtol=1e-12, equal_nan=True)


def _pair_stats_equal(got, want):
    assert sorted(got) == sorted(want)
    for k, v in want.items():
        if k == "unique":
            assert sorted(got[k]) == sorted(v)
            for name in v:
                assert got[k][name][1] == v[name][1]
                assert got[k][name][0] == v[name][0] or (v[name][0] is None and np.isnan(got[k][name][0]))
        else:
            assert got[k][1] == v[1], k
            assert got[k][0] == v[0] or (v[0] is None and np.isnan(got[k][0])), k


def test_pairsnp_cases(small_panel):
    """pairwiseScore (snpmatch.py:270-309) of the reference, with and without a database."""
    import json
    import os
    from conftest import GOLDEN
    g = load_golden("pairsnp.npz")
    with open(os.path.join(GOLDEN, "pairsnp.json")) as fh:
        want = json.load(fh)
    db = (orc.db_chromosome_labels(small_panel["chrs"], small_panel["chr_regions"]), small_panel["positions"])
    for i in range(int(g["n_cases"])):
        a = [g["p%d_%s" % (i, k)] for k in ("c1", "p1", "g1", "c2", "p2", "g2")]
        names = ("s%d_a.npz" % i, "s%d_b.npz" % i)
        _pair_stats_equal(orc.pairwise_score(*a, name1=names[0], name2=names[1]), want["p%d_plain" % i])
        _pair_stats_equal(orc.pairwise_score(*a, name1=names[0], name2=names[1], db=db), want["p%d_db" % i])


def test_simulate_cases(small_panel):
    """simulateSNPs / simulateSNPs_F1 (simulate.py:10-60) under the reference's np.random call order."""
    g = load_golden("simulate.npz")
    p = small_panel
    ids = p["accessions"].astype("U")
    row_chrs = np.array(orc.db_chromosome_labels(p["chrs"], p["chr_regions"]))
    for i in range(int(g["n_cases"])):
        n, err, rm, seed = g["s%d_args" % i]
        np.random.seed(int(seed))
        if str(g["s%d_kind" % i]) == "inbred":
            col = p["snps"][:, np.flatnonzero(ids == str(g["s%d_acc" % i]))[0]]
            c, pos, snp = orc.simulate_snps(col, row_chrs, p["positions"], int(n), float(err))
        else:
            a, b = str(g["s%d_acc" % i]).split("x")
            c, pos, snp = orc.simulate_snps_f1(p["snps"][:, np.flatnonzero(ids == a)[0]], p["snps"][:, np.flatnonzero(ids == b)[0]],
                                               row_chrs, p["positions"], int(n), float(err), float(rm))
        gt = np.array(["./.", "0/0", "1/1", "0/1"])[snp.astype(int) + 1]
        assert np.array_equal(c.astype("U"), g["s%d_chr" % i])
        assert np.array_equal(pos, g["s%d_pos" % i])
        assert np.array_equal(gt, g["s%d_gt" % i])


def _gc_lines(calls, samples, genome_chrs, genome_chrlen, rates, bin_len):
    """Output lines of genotype_cross (genotype_cross.py:217-237) from the oracle's per-window calls."""
    ids = orc.genome_chr_ids(genome_chrs)
    lines = ["id,,," + ",".join(samples), "pheno,," + ",0" * len(samples)]
    w = 0
    for ci, cid in enumerate(ids):
        for k in range(orc.num_windows(genome_chrlen[ci], bin_len)):
            start, end = 1 + k * bin_len, (k + 1) * bin_len
            cm = rates[ci] * int(round(float(np.mean([start, end])))) / 1000000
            tail = ",NA" * len(samples) if calls[w] is None else "".join("," + str(g) for g in calls[w])
            lines.append("%s:%d-%d,%s,%s%s" % (cid, start, end, cid, cm, tail))
            w += 1
    return lines


TAIR10 = (["1", "2", "3", "4", "5"], [30427671, 19698289, 23459830, 18585056, 26975502], [3.4, 3.6, 3.5, 3.8, 3.6])


def test_genotype_cross_cases(small_panel):
    """GenotypeCross.genotype_cross and getWindowGenotype (genotype_cross.py:21-49,210-241) of the reference."""
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "genotype_cross.json")) as fh:
        want = json.load(fh)
    for total, a, h, b, lr, geno in want["window_genotype_grid"]:
        got = orc.get_window_genotype([a, h, b], total, lr)
        assert (-1 if got == "NA" else got) == geno, (total, a, h, b, lr)
    vcf = load_golden("genotype_cross_vcf.npz")
    p = small_panel
    ids = p["accessions"].astype("U")
    row_chrs = np.array(orc.db_chromosome_labels(p["chrs"], p["chr_regions"]))
    for tag in ("b300k_lr1.5", "b1M_lr3", "b2M_lr1.5"):
        c = want[tag]
        i1, i2 = [int(np.flatnonzero(ids == x)[0]) for x in c["parents"].split("x")]
        seg = orc.segregating_parent_markers(p["snps"][:, i1], p["snps"][:, i2])
        assert len(seg) == c["n_segregating"]
        calls, _, _ = orc.genotype_cross_windows(row_chrs[seg], p["positions"][seg], p["snps"][seg, i1], p["snps"][seg, i2],
                                                 vcf["chr"], vcf["pos"], vcf["gt"], TAIR10[0], TAIR10[1], c["bin_len"], c["lr_thres"])
        assert _gc_lines(calls, vcf["samples"], TAIR10[0], TAIR10[1], TAIR10[2], c["bin_len"]) == c["lines"]
