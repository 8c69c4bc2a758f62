"""SURVEY 8(f)-3 on the GPU: column reads of the resident panel, `pairsnp` and `simulate`, against the CPU oracle and the
golden vectors generated from the unmodified reference (tests/golden/pairsnp.*, simulate.npz).  Everything is integer /
string work: all comparisons are exact."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import snpmatch_oracle as orc
from snpmatch_b200 import synth

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("n_acc", [1, 31, 33, 64, 300, 1135])
def test_read_columns(lib, n_acc):
    rng = np.random.default_rng(100 + n_acc)
    n = 1031
    snps = rng.choice(np.array([-1, 0, 1, 2], dtype=np.int8), size=(n, n_acc), p=[0.1, 0.5, 0.3, 0.1])
    db = lib.Database(np.arange(1, n + 1, dtype=np.int32), np.array([[0, n]]), n_acc)
    db.load_int8(snps)
    for k in (1, 2, 16, 17, 40):                     # more than one launch of 16 columns, repeated columns allowed
        cols = rng.integers(0, n_acc, size=k)
        assert np.array_equal(db.read_columns(cols), snps[:, cols].T)
    assert db.read_columns(np.zeros(0, dtype=np.int32)).shape == (0, n)
    with pytest.raises(lib.SnpmError):
        db.read_columns([n_acc])
    db.close()


def test_snps_view_columns(lib, small_geno, small_panel):
    snps = small_panel["snps"]
    v = small_geno.g_acc.snps
    assert np.array_equal(v[:, 7], snps[:, 7])
    assert np.array_equal(v[:, [3, 21]], snps[:, [3, 21]])
    assert np.array_equal(v[:, -1], snps[:, -1])
    assert np.array_equal(v[[5, 9], :], snps[[5, 9], :])          # the row path is untouched


def test_read_columns_full_size_panel(lib):
    """10.7 M rows x 1135 accessions generated in HBM: columns against the host hash on sampled rows, and called-genotype
    counts of a whole column against the per-class totals of the same rows unpacked row-wise."""
    from snpmatch_b200.core import snp_genotype
    g = snp_genotype.Genotype.synthetic(10_700_000, 1135)
    cols = np.array([0, 7, 31, 32, 1134])
    out = g.g.db.read_columns(cols)
    assert out.shape == (5, 10_700_000)
    rows = np.random.default_rng(5).integers(0, 10_700_000, size=20000)
    assert np.array_equal(out[:, rows].T, synth.panel_codes_cols(synth.SEED_PANEL, rows, cols))
    assert np.array_equal(out[:, rows].T, g.g.db.read_rows(rows)[:, cols])
    assert set(np.unique(out).tolist()) <= {-1, 0, 1, 2}
    g.close()


def test_pair_match_counts_vs_numpy(lib):
    rng = np.random.default_rng(77)
    for n1, n2, m, n_chr in ((5000, 4000, 3000, 5), (100, 100, 0, 3), (20000, 20000, 20000, 300), (10, 10, 10, 1)):
        chrom1 = rng.integers(-1, n_chr + 1, size=n1)             # ids outside 0..n_chr-1 are not counted
        gt1, gt2 = rng.integers(0, 4, size=n1), rng.integers(0, 4, size=n2)
        i1, i2 = rng.integers(0, n1, size=m), rng.integers(0, n2, size=m)
        common, matches = lib.pair_match_counts(i1, i2, chrom1, gt1, gt2, n_chr)
        c = chrom1[i1]
        ok = (c >= 0) & (c < n_chr)
        assert np.array_equal(common, np.bincount(c[ok], minlength=n_chr))
        assert np.array_equal(matches, np.bincount(c[ok & (gt1[i1] == gt2[i2])], minlength=n_chr))
    with pytest.raises(lib.SnpmError):
        lib.pair_match_counts([5], [0], [0], [0], [0], 1)


def _stats_equal(got, want):
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


def test_pairwise_score_golden(lib, small_geno, tmp_path):
    from snpmatch_b200.core import snpmatch
    g = load_golden("pairsnp.npz")
    with open(os.path.join(GOLDEN, "pairsnp.json")) as fh:
        want = json.load(fh)
    for i in range(int(g["n_cases"])):
        files = []
        for side, tag in (("1", "a"), ("2", "b")):
            f = str(tmp_path / ("s%d_%s.npz" % (i, tag)))
            n = len(g["p%d_p%s" % (i, side)])
            np.savez(f, chr=g["p%d_c%s" % (i, side)], pos=g["p%d_p%s" % (i, side)], gt=g["p%d_g%s" % (i, side)], wei=np.ones((n, 3)), dp=np.ones(n))
            files.append(f)
        _stats_equal(snpmatch.pairwiseScore(files[0], files[1], False, outFile=None, hdf5File=None), want["p%d_plain" % i])
        r = snpmatch.pairwiseScore(files[0], files[1], False, outFile=str(tmp_path / ("o%d" % i)), hdf5File=small_geno)
        r.pop("hdf5")
        _stats_equal(r, want["p%d_db" % i])
        on_disk = json.load(open(str(tmp_path / ("o%d" % i)) + ".matches.json"))     # the reference cannot write this file under Python 3
        assert on_disk["matches"][1] == want["p%d_db" % i]["matches"][1]


def test_simulate_golden(lib, small_geno, tmp_path):
    from snpmatch_b200.core import simulate
    g = load_golden("simulate.npz")
    for i in range(int(g["n_cases"])):
        n, err, rm, seed = g["s%d_args" % i]
        np.random.seed(int(seed))
        out = str(tmp_path / ("sim%d.bed" % i))
        if str(g["s%d_kind" % i]) == "inbred":
            df = simulate.simulateSNPs(small_geno, str(g["s%d_acc" % i]), int(n), outFile=out, err_rate=float(err))
        else:
            df = simulate.simulateSNPs_F1(small_geno, str(g["s%d_acc" % i]), int(n), out, float(err), float(rm))
        assert np.array_equal(np.array(df["chr"]).astype("U"), g["s%d_chr" % i])
        assert np.array_equal(np.array(df["pos"]).astype(np.int64), g["s%d_pos" % i])
        assert np.array_equal(np.array(df["snp"]).astype("U"), g["s%d_gt" % i])
        lines = open(out).read().strip("\n").split("\n")
        assert lines[0].split("\t") == [str(g["s%d_chr" % i][0]), str(g["s%d_pos" % i][0]), str(g["s%d_gt" % i][0])]
        assert len(lines) == int(n)
    with pytest.raises(AssertionError):
        simulate.simulateSNPs(small_geno, "no-such-accession", 10)


def test_simulated_sample_is_genotyped_back(lib, small_geno, tmp_path):
    """simulate -> inbred round trip (what `snpmatch simulate` exists for): the accession the markers were drawn from is the top hit."""
    from snpmatch_b200.core import parsers, simulate, snpmatch
    np.random.seed(3)
    bed = str(tmp_path / "sim.bed")
    simulate.simulateSNPs(small_geno, str(small_geno.accessions[13]), 1500, outFile=bed, err_rate=0.01)
    inputs = parsers.ParseInputs(inFile=bed, logDebug=False)
    gt = snpmatch.Genotyper(inputs, small_geno, str(tmp_path / "rt"), run_genotyper=True)
    assert int(np.nanargmin(gt.result.likelis)) == 13
    assert gt.result.num_snps == 1500


def test_command_line_pairsnp_and_simulate(lib, small_geno, tmp_path):
    import snpmatch_b200
    db_path = str(tmp_path / "panel.npz")
    small_geno.save_packed(db_path)
    a, b = str(tmp_path / "a.bed"), str(tmp_path / "b.bed")
    ids = small_geno.accessions
    np.random.seed(1)
    assert snpmatch_b200.main(["simulate", "-d", db_path, "-a", str(ids[3]), "-n", "800", "-p", "0.0", "-o", a]) == 0
    assert snpmatch_b200.main(["simulate", "-d", db_path, "-a", "%sx%s" % (ids[3], ids[21]), "--f1", "-n", "900", "-p", "0.0", "-o", b]) == 0
    out = str(tmp_path / "pair")
    assert snpmatch_b200.main(["pairsnp", "-i", a, "-j", b, "-d", db_path, "-o", out]) == 0
    js = json.load(open(out + ".matches.json"))
    ca, cb = np.loadtxt(a, dtype=str), np.loadtxt(b, dtype=str)
    st = orc.pairwise_score(ca[:, 0], ca[:, 1].astype(int), ca[:, 2], cb[:, 0], cb[:, 1].astype(int), cb[:, 2], "a.bed", "b.bed")
    assert js["matches"] == st["matches"]
    for c in ("1", "2", "3", "4", "5"):
        assert js[c] == st[c]


def test_makedb_builds_a_loadable_database(lib, small_panel, sample_inbred, tmp_path):
    """CSV (the reference's intermediate makedb format) -> streamed to the GPU in small chunks -> packed .npz -> loaded again:
    same matrix, same index arrays; `inbred` on it gives the scores of the panel built from arrays; and the command line."""
    import snpmatch_b200
    from snpmatch_b200.core import makedb, snp_genotype, parsers, snpmatch
    p = small_panel
    labels = np.array(orc.db_chromosome_labels(p["chrs"], p["chr_regions"]))
    csv = str(tmp_path / "panel.csv")
    with open(csv, "w") as fh:
        fh.write("Chromosome,Position," + ",".join(p["accessions"].astype("U")) + "\n")
        for c, pos, row in zip(labels, p["positions"], p["snps"]):
            fh.write("%s,%d,%s\n" % (c, pos, ",".join(str(int(v)) for v in row)))
    g = makedb.makeDB(csv, str(tmp_path / "db"), chunk_rows=1700)          # 6000 rows in four chunks
    assert np.array_equal(g.g.snps[:, :], p["snps"])
    g.close()
    g2 = snp_genotype.Genotype(str(tmp_path / "db.npz"))
    assert np.array_equal(g2.g.snps[:, :], p["snps"]) and np.array_equal(g2.g.positions, p["positions"])
    assert g2.chrs.tolist() == p["chrs"].astype("U").tolist() and np.array_equal(g2.g.chr_regions, p["chr_regions"])
    assert g2.accessions.tolist() == p["accessions"].astype("U").tolist()
    s = sample_inbred
    inp = parsers.ParseInputs("")
    inp.load_snp_info(s["chrs"], s["pos"], s["gt"], s["wei"], s["dp"])
    r = snpmatch.Genotyper(inp, g2, str(tmp_path / "o"), run_genotyper=True).result
    want = load_golden("inbred_pl.npz")
    assert np.array_equal(r.scores, want["scores"]) and np.array_equal(r.ninfo, want["ninfo"])
    g2.close()
    assert snpmatch_b200.main(["makedb", "-i", csv, "-o", str(tmp_path / "cli_db")]) == 0
    g3 = snp_genotype.Genotype(str(tmp_path / "cli_db.npz"))
    assert np.array_equal(g3.g.snps[[0, 17, 5999], :], p["snps"][[0, 17, 5999], :])
    g3.close()
