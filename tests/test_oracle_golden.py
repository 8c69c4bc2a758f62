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
