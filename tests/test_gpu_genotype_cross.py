"""SURVEY 8(f)-4 on the GPU: `genotype_cross` window calls against the golden output of the unmodified reference
(tests/golden/genotype_cross.json) and the CPU oracle.  Counts and calls are integers: compared exactly."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import snpmatch_oracle as orc

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


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "genotype_cross.json")) as fh:
        return json.load(fh)


def _write_vcf(path, vcf):
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n")
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(vcf["samples"]) + "\n")
        for c, p, row in zip(vcf["chr"], vcf["pos"], vcf["gt"]):
            fh.write("%s\t%d\t.\tA\tT\t50\tPASS\t.\tGT\t%s\n" % (c, p, "\t".join(row)))


def test_window_genotype_grid(lib, golden):
    """getWindowGenotype (genotype_cross.py:21-49) on the device: one window per grid cell, one sample."""
    from snpmatch_b200.core import genotype_cross as gc
    grid = np.array(golden["window_genotype_grid"])
    for lr in (1.5, 3.0):
        cells = grid[grid[:, 4] == lr]
        # window k: `total` markers; the sample equals parent 1 on the first a, is het on the next h, equals parent 2 on the next b
        p1, p2, gt, ws = [], [], [], [0]
        for total, a, h, b, _, _ in cells.astype(int):
            p1 += [0] * total
            p2 += [1] * total
            gt += [0] * a + [2] * h + [1] * b + [-1] * (total - a - h - b)
            ws.append(ws[-1] + total)
        ix = np.arange(ws[-1])
        counts, geno, border = lib.cross_window_genotypes(ix, ix, ws, p1, p2, np.array(gt, dtype=np.int8).reshape(-1, 1), lr)
        assert np.array_equal(counts[:, 0, :], cells[:, [1, 2, 3]].astype(np.int32))
        assert np.array_equal(geno[:, 0], cells[:, 5].astype(np.int8))
        assert not border.any()
    for total, a, h, b, lr, want in grid[::37]:
        geno, pval = gc.getWindowGenotype([int(a), int(h), int(b)], int(total), lr)
        assert (-1 if geno == "NA" else geno) == want


@pytest.mark.parametrize("tag", ["b300k_lr1.5", "b1M_lr3", "b2M_lr1.5"])
def test_genotype_cross_golden(lib, small_geno, golden, tmp_path, tag):
    from snpmatch_b200.core import genotype_cross as gc
    c = golden[tag]
    vcf = load_golden("genotype_cross_vcf.npz")
    path = str(tmp_path / "population.vcf")
    _write_vcf(path, vcf)
    x = gc.GenotypeCross(small_geno, c["parents"], c["bin_len"], None, False, genome_id="athaliana_tair10")
    assert len(x.commonSNPsPOS) == c["n_segregating"]
    lines = x.genotype_cross(path, c["lr_thres"])
    assert [str(l) for l in lines] == c["lines"]
    assert not x.last_window_calls["borderline"].any()
    # counts against the oracle
    p = {k: small_geno.g.snps[:, :][:, i] for k, i in (("p1", x.p1_ix), ("p2", x.p2_ix))}
    seg = orc.segregating_parent_markers(p["p1"], p["p2"])
    _, counts, n_matched = orc.genotype_cross_windows(x.commonSNPsCHR, x.commonSNPsPOS, p["p1"][seg], p["p2"][seg], vcf["chr"], vcf["pos"],
                                                      vcf["gt"], ["1", "2", "3", "4", "5"], x.genome.chrlen, c["bin_len"], c["lr_thres"])
    assert np.array_equal(x.last_window_calls["n_matched"], n_matched)
    for w, cnt in counts.items():
        assert np.array_equal(x.last_window_calls["counts"][w], cnt)
    out = str(tmp_path / "gc.csv")
    x.write_output_genotype_cross(lines, out)
    assert open(out).read().split("\n")[:-1] == c["lines"]


def test_genotype_cross_parents_from_files_and_cli(lib, small_geno, golden, tmp_path):
    """Parents given as two files that list the same positions (the one layout the reference's --father mode supports) call the
    same windows as the database parents restricted to those positions; and the command line end to end."""
    import snpmatch_b200
    from snpmatch_b200.core import genotype_cross as gc
    c = golden["b1M_lr3"]
    vcf = load_golden("genotype_cross_vcf.npz")
    path = str(tmp_path / "population.vcf")
    _write_vcf(path, vcf)
    db_path = str(tmp_path / "panel.npz")
    small_geno.save_packed(db_path)
    out = str(tmp_path / "cli.csv")
    assert snpmatch_b200.main(["genotype_cross", "-i", path, "-d", db_path, "-p", c["parents"], "-b", str(c["bin_len"]), "--lr_thres",
                               str(c["lr_thres"]), "-o", out]) == 0
    assert open(out).read().split("\n")[:-1] == c["lines"]
    # parents as BED files over all panel positions
    ids = small_geno.accessions
    i1, i2 = [int(np.flatnonzero(ids == x)[0]) for x in c["parents"].split("x")]
    cols = small_geno.g_acc.snps[:, [i1, i2]]
    chrom = np.array(small_geno.g.chromosomes)
    names = np.array(["./.", "0/0", "1/1", "0/1"])
    for tag, col in (("mother", cols[:, 0]), ("father", cols[:, 1])):
        with open(str(tmp_path / (tag + ".bed")), "w") as fh:
            for ch, p, g in zip(chrom, small_geno.g.positions, names[col.astype(int) + 1]):
                fh.write("%s\t%d\t%s\n" % (ch, p, g))
    x = gc.GenotypeCross(small_geno, str(tmp_path / "mother.bed"), c["bin_len"], str(tmp_path / "father.bed"), False, genome_id="athaliana_tair10")
    assert len(x.commonSNPsPOS) == c["n_segregating"]
    assert [str(l) for l in x.genotype_cross(path, c["lr_thres"])] == c["lines"]
    with pytest.raises(NotImplementedError):
        x.genotype_cross_hmm(path)


def test_window_calls_many_samples_vs_oracle(lib):
    """300 samples (three sample tiles), ragged windows, empty windows, positions outside every window."""
    from snpmatch_b200.core import genomes, genotype_cross as gc
    rng = np.random.default_rng(9)
    gen = genomes.Genome("athaliana_tair10")
    n_par, n_vcf, S = 4000, 5000, 300
    def markers(n):
        c = np.sort(rng.integers(1, 6, size=n))
        p = np.concatenate([np.sort(rng.choice(31_000_000, size=int((c == k).sum()), replace=False)) + 1 for k in range(1, 6)])
        return np.char.add("Chr", c.astype(str)), p
    pc, pp = markers(n_par)
    hole = (pc == "Chr3") & (pp > 5_000_000) & (pp <= 7_000_000)          # four windows without parental markers
    pc, pp = pc[~hole], pp[~hole]
    n_par = len(pp)
    vc, vp = markers(n_vcf)
    share = rng.choice(n_par, 2500, replace=False)           # force common positions
    vc, vp = np.concatenate([vc, pc[share]]), np.concatenate([vp, pp[share]])
    o = np.lexsort((vp, vc))
    vc, vp = vc[o], vp[o]
    keep = np.ones(len(vp), dtype=bool)
    keep[1:] = ~((vc[1:] == vc[:-1]) & (vp[1:] == vp[:-1]))
    vc, vp = vc[keep], vp[keep]
    p1 = rng.integers(0, 2, size=n_par).astype(np.int8)
    p2 = (1 - p1).astype(np.int8)
    codes = rng.choice(np.array([-1, 0, 1, 2], dtype=np.int8), size=(len(vp), S), p=[0.1, 0.4, 0.3, 0.2])
    r = gc.window_calls(pc, pp, p1, p2, vc, vp, codes, gen, 500000, 2.0)
    names = np.array(["./.", "0/0", "1/1", "0/1"])
    calls, counts, n_matched = orc.genotype_cross_windows(pc, pp, p1, p2, vc, vp, names[codes[:, ::29].astype(int) + 1], gen.chrs, gen.chrlen, 500000, 2.0)
    assert np.array_equal(r["n_matched"], n_matched)
    assert (r["n_matched"] == 0).any() and (r["n_matched"] > 5).any()
    for w, cnt in counts.items():
        assert np.array_equal(r["counts"][w][::29], cnt)
        assert [(-1 if g == "NA" else g) for g in calls[w]] == r["geno"][w][::29].tolist()
    assert (r["geno"][r["n_matched"] == 0] == -1).all()
