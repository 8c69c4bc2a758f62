"""
TEST INFRASTRUCTURE — golden-vector generator.

Runs the UNMODIFIED reference (imported read-only from /root/reference through
oracle/ref_harness.py) on seeded inputs and writes the results to tests/golden/.
Run in the build container only:   python -m oracle.gen_golden
The committed vectors are what pins both the oracle restatement
(tests/test_oracle_golden.py) and, on the GPU box, the CUDA path.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh            # noqa: E402
from snpmatch_b200 import synth                 # noqa: E402
from snpmatch_b200.core import parsers as my_parsers   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _pl_like_weights(rng, k):
    dp = 1 + rng.poisson(3, size=k)
    code = rng.choice([0, 1, 2], size=k, p=[0.7, 0.25, 0.05]).astype(np.int8)
    return synth._pl_weights(rng, code, dp)[1], code


def gen_match_cases(ref):
    rng = np.random.default_rng(11)
    cases = {}
    shapes = [(1, 1), (7, 5), (64, 33), (300, 70), (1000, 129), (257, 300)]
    ci = 0
    for (k, a) in shapes:
        for kind in ("pl", "hard"):
            for skip in (False, True):
                snps = rng.choice(np.array([-1, 0, 1, 2], dtype=np.int8), size=(k, a), p=[0.1, 0.55, 0.3, 0.05])
                if a > 3:
                    snps[:, 1] = -1          # an all-missing accession -> ninfo 0
                    snps[:, 2] = 0           # all ref
                wei, code = _pl_like_weights(rng, k)
                if kind == "hard":
                    wei = synth.hard_weights(code)
                if a > 4:
                    # an accession that matches the sample exactly where it is called hom
                    snps[:, 3] = np.where(code == 2, 2, code)
                score, ninfo = ref.snpmatch.matchGTsAccs(wei.copy(), snps.copy(), skip)
                cases["c%d_snps" % ci] = snps
                cases["c%d_wei" % ci] = wei
                cases["c%d_skip" % ci] = np.array(skip)
                cases["c%d_score" % ci] = np.asarray(score, dtype=np.float64)
                cases["c%d_ninfo" % ci] = np.asarray(ninfo, dtype=np.int64)
                ci += 1
    cases["n_cases"] = np.array(ci)
    np.savez_compressed(os.path.join(GOLD, "match_gts_accs.npz"), **cases)
    print("match_gts_accs: %d cases" % ci)


def gen_join_cases(ref):
    rng = np.random.default_rng(12)
    out = {}
    cases = []
    # the probed example of SURVEY A.1
    cases.append((np.array(['1'] * 5 + ['2'] * 4), np.array([10, 20, 30, 40, 50, 5, 15, 25, 35]),
                  np.array(['Chr2'] * 3 + ['Chr1'] * 3), np.array([15, 16, 35, 20, 50, 60])))
    # prefixes, extra contigs, chromosomes missing on either side
    cases.append((np.array(['Chr1'] * 4 + ['Chr3'] * 3 + ['chrC'] * 2), np.array([3, 9, 12, 40, 1, 2, 3, 7, 8]),
                  np.array(['1'] * 3 + ['2'] * 2 + ['ChrM'] * 2 + ['C'] * 2), np.array([9, 12, 13, 1, 2, 5, 6, 8, 9])))
    # disjoint
    cases.append((np.array(['1'] * 3), np.array([1, 2, 3]), np.array(['1'] * 2), np.array([7, 8])))
    # random larger
    for _ in range(3):
        n1, n2 = 4000, 700
        c1 = np.sort(rng.integers(1, 6, size=n1))
        p1 = np.concatenate([np.sort(rng.choice(50000, size=(c1 == c).sum(), replace=False)) + 1 for c in range(1, 6)])
        c2 = np.sort(rng.integers(1, 7, size=n2))
        p2 = np.concatenate([np.sort(rng.choice(50000, size=(c2 == c).sum(), replace=False)) + 1 for c in range(1, 7)])
        cases.append((c1.astype(str), p1, np.char.add("Chr", c2.astype(str)), p2))
    for i, (c1, p1, c2, p2) in enumerate(cases):
        i1, i2 = ref.snp_genotype.Genotype.get_common_positions(c1, p1, c2, p2)
        out["j%d_c1" % i] = np.asarray(c1, dtype="str")
        out["j%d_p1" % i] = np.asarray(p1, dtype=np.int64)
        out["j%d_c2" % i] = np.asarray(c2, dtype="str")
        out["j%d_p2" % i] = np.asarray(p2, dtype=np.int64)
        out["j%d_i1" % i] = np.asarray(i1, dtype=np.int64)
        out["j%d_i2" % i] = np.asarray(i2, dtype=np.int64)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, "join.npz"), **out)
    print("join: %d cases" % len(cases))


def gen_epilogue_cases(ref):
    rng = np.random.default_rng(13)
    out = {}
    sets = []
    sets.append((np.array([0, 0, 3, 10, 5.5]), np.array([0, 10, 10, 10, 10])))
    n = rng.integers(0, 3000, size=400)
    y = np.floor(n * rng.random(400) ** 0.2).astype(np.int64)
    y[::7] = n[::7]
    y[::11] = 0
    sets.append((y, n))
    nf = rng.integers(1, 400, size=300)
    yf = nf * rng.random(300)
    yf[::5] = nf[::5]
    sets.append((yf, nf))
    sets.append((np.zeros(5), np.array([0, 1, 2, 3, 4])))         # all nan
    sets.append((np.array([5, 7]), np.array([5, 7])))              # all perfect
    for i, (y, n) in enumerate(sets):
        L, LR = ref.snpmatch.GenotyperOutput.calculate_likelihoods(y, n)
        out["e%d_y" % i] = np.asarray(y, dtype=np.float64)
        out["e%d_n" % i] = np.asarray(n, dtype=np.int64)
        out["e%d_L" % i] = np.asarray(L, dtype=np.float64)
        out["e%d_LR" % i] = np.asarray(LR, dtype=np.float64)
    out["n_sets"] = np.array(len(sets))
    # identity test grid (csmatch default error rate 0.02 and snpmatch default 0.0005)
    ns, xs = [], []
    for nn in list(range(0, 60)) + [100, 200, 500, 1000, 2000]:
        for miss in range(0, min(nn, 40) + 1):
            ns.append(nn)
            xs.append(nn - miss)
    ns = np.array(ns)
    xs = np.array(xs, dtype=np.float64)
    xs_frac = xs - (np.arange(len(xs)) % 3) * 0.25
    xs_frac = np.clip(xs_frac, 0, None)
    out["id_n"] = ns
    out["id_x"] = xs
    out["id_xf"] = xs_frac
    out["id_e02"] = ref.snpmatch.np_test_identity(xs, ns, error_rate=0.02)
    out["id_e02_f"] = ref.snpmatch.np_test_identity(xs_frac, ns, error_rate=0.02)
    out["id_default"] = ref.snpmatch.np_test_identity(xs, ns)
    out["likeli_10_3"] = np.array(ref.snpmatch.likeliTest(10, 3))
    np.savez_compressed(os.path.join(GOLD, "epilogue.npz"), **out)
    print("epilogue: %d sets, %d identity points" % (len(sets), len(ns)))


def _run_reference_inbred(ref, panel, sample, wei, out_prefix, skip):
    G = rh.make_reference_genotype(ref, panel["snps"], panel["positions"], panel["chrs"],
                                   panel["chr_regions"], panel["accessions"])
    inp = rh.make_reference_inputs(ref, sample["chrs"], sample["pos"], sample["gt"], wei, sample["dp"])
    gt = ref.snpmatch.Genotyper(inp, G, out_prefix, run_genotyper=True, skip_db_hets=skip)
    return gt


def _run_reference_cross(ref, panel, sample, wei, out_prefix, skip, bin_len=300000):
    G = rh.make_reference_genotype(ref, panel["snps"], panel["positions"], panel["chrs"],
                                   panel["chr_regions"], panel["accessions"])
    inp = rh.make_reference_inputs(ref, sample["chrs"], sample["pos"], sample["gt"], wei, sample["dp"])
    ci = ref.csmatch.CrossIdentifier(inp, G, "athaliana_tair10", bin_len, out_prefix, run_identifier=True,
                                     skip_db_hets=skip)
    return ci


def _read(path):
    with open(path) as fh:
        return fh.read()


def gen_workflow_cases(ref):
    """Whole `inbred` and `cross` runs of the reference on a small synthetic panel."""
    panel = synth.small_panel(n_rows=6000, n_acc=40)
    # make accession 9 a near-duplicate of accession 7 (an ambiguous pair), accession 11 all-missing
    panel["snps"][:, 9] = panel["snps"][:, 7]
    flip = np.arange(0, 6000, 37)
    panel["snps"][flip, 9] = np.where(panel["snps"][flip, 9] == 0, 1, 0)
    panel["snps"][:, 11] = -1
    np.savez_compressed(os.path.join(GOLD, "small_panel.npz"), **panel)
    tmp = tempfile.mkdtemp(prefix="snpm_golden_")
    index = {}
    try:
        # ---- inbred
        s_in = synth.make_sample(panel["positions"], panel["chr_regions"], panel["chrs"], 40, true_acc=7,
                                 n_db=2600, n_extra=300, seed=501)
        np.savez_compressed(os.path.join(GOLD, "sample_inbred.npz"),
                            **{k: s_in[k] for k in ("chrs", "pos", "gt", "wei", "wei_hard", "dp")})
        for tag, wei_key, skip in (("pl", "wei", False), ("pl_skip", "wei", True), ("hard", "wei_hard", False)):
            pre = os.path.join(tmp, "inbred_" + tag)
            gt = _run_reference_inbred(ref, panel, s_in, s_in[wei_key], pre, skip)
            r = gt.result
            np.savez_compressed(os.path.join(GOLD, "inbred_%s.npz" % tag),
                                scores=r.scores, ninfo=r.ninfo, likelis=r.likelis, lrts=r.lrts,
                                probs=r.probabilies, overlap=np.array(r.overlap), num_snps=np.array(r.num_snps),
                                common_db=np.asarray(gt.commonSNPs[0]), common_s=np.asarray(gt.commonSNPs[1]))
            index["inbred_" + tag] = {"scores.txt": _read(pre + ".scores.txt"),
                                      "matches.json": _read(pre + ".matches.json")}
        # ---- cross (F2-like mosaic of accessions 3 and 21)
        s_cr = synth.make_sample(panel["positions"], panel["chr_regions"], panel["chrs"], 40, n_db=3000, n_extra=300,
                                 seed=777, mosaic=(3, 21, 3000000), err=0.002, het=0.0)
        np.savez_compressed(os.path.join(GOLD, "sample_cross.npz"),
                            **{k: s_cr[k] for k in ("chrs", "pos", "gt", "wei", "wei_hard", "dp")})
        for tag, sample, wei_key, skip in (("pl", s_cr, "wei", False), ("hard_skip", s_cr, "wei_hard", True),
                                           ("inbredlike", s_in, "wei", False)):
            pre = os.path.join(tmp, "cross_" + tag)
            ci = _run_reference_cross(ref, panel, sample, sample[wei_key], pre, skip)
            r = ci.result
            files = {}
            for suffix in (".windowscore.txt", ".scores.txt", ".scores.txt.matches.json", ".matches.json"):
                if os.path.exists(pre + suffix):
                    files[suffix[1:]] = _read(pre + suffix)
            index["cross_" + tag] = files
            np.savez_compressed(os.path.join(GOLD, "cross_%s.npz" % tag),
                                scores=np.asarray(r.scores, dtype=np.float64), ninfo=np.asarray(r.ninfo, dtype=np.int64),
                                accs=np.asarray(r.accs, dtype="str"), likelis=r.likelis, lrts=r.lrts,
                                num_snps=np.array(r.num_snps), overlap=np.array(r.overlap),
                                matchedTarInd=np.asarray(r.matchedTarInd, dtype=np.int64),
                                winds_chrs=np.asarray(r.winds_chrs, dtype="str"))
        # ---- repo-data config: the shipped sample VCF against a synthetic DB over its positions
        vcf = os.path.join(rh.REFERENCE_ROOT, "sample_files", "701_501.filter.vcf")
        scratch = os.path.join(tmp, "701_501.filter.vcf")
        shutil.copy(vcf, scratch)                       # never parse inside /root/reference (writes a cache)
        pi = my_parsers.ParseInputs("")
        pi.load_snp_info(*pi.read_vcf(scratch, True))
        raw = my_parsers.read_vcf_minimal(scratch)
        rng = np.random.default_rng(701)
        # DB rows: every VCF record position (kept or not), per chromosome in file order
        chrs_all = np.array([c.replace("Chr", "") for c in raw["chr"]])
        db_chrs = np.array(["1", "2", "3", "4", "5"])
        regions, positions = [], []
        start = 0
        for c in db_chrs:
            p = raw["pos"][chrs_all == c]
            positions.append(p)
            regions.append((start, start + len(p)))
            start += len(p)
        positions = np.concatenate(positions).astype(np.int32)
        n_acc = 60
        freq = rng.random(len(positions)) ** 3
        snps = (rng.random((len(positions), n_acc)) < freq[:, None]).astype(np.int8)
        snps[rng.random(snps.shape) < 0.05] = -1
        snps[rng.random(snps.shape) < 0.002] = 2
        # accession 17 carries the sample's own calls
        code_all = my_parsers.parseGT(raw["gt"])
        code_all = np.where(code_all < 0, 0, code_all)
        order = np.concatenate([np.flatnonzero(chrs_all == c) for c in db_chrs])
        snps[:, 17] = code_all[order]
        noise = rng.random(len(positions)) < 0.01
        snps[noise, 17] = -1
        panel_vcf = dict(snps=snps, positions=positions, chr_regions=np.array(regions, dtype=np.int64),
                         chrs=db_chrs, accessions=synth.accession_ids(n_acc))
        np.savez_compressed(os.path.join(GOLD, "vcf701_panel.npz"), **panel_vcf)
        dp = np.asarray(pi.dp, dtype=np.float64)
        np.savez_compressed(os.path.join(GOLD, "vcf701_sample.npz"), chrs=pi.chrs, pos=pi.pos, gt=pi.gt, wei=pi.wei, dp=dp)
        s_vcf = dict(chrs=pi.chrs, pos=pi.pos, gt=pi.gt, wei=pi.wei, dp=dp)
        pre = os.path.join(tmp, "vcf701_inbred")
        gt = _run_reference_inbred(ref, panel_vcf, s_vcf, pi.wei, pre, False)
        index["vcf701_inbred"] = {"scores.txt": _read(pre + ".scores.txt"), "matches.json": _read(pre + ".matches.json")}
        pre = os.path.join(tmp, "vcf701_cross")
        _run_reference_cross(ref, panel_vcf, s_vcf, pi.wei, pre, False)
        files = {}
        for suffix in (".windowscore.txt", ".scores.txt", ".scores.txt.matches.json", ".matches.json"):
            if os.path.exists(pre + suffix):
                files[suffix[1:]] = _read(pre + suffix)
        index["vcf701_cross"] = files
        facts = {"vcf_kept": int(len(pi.chrs)), "vcf_wei_colsum": pi.wei.sum(axis=0).tolist(),
                 "vcf_dp_mean": float(np.mean(dp)), "vcf_first": [str(pi.chrs[0]), int(pi.pos[0]), str(pi.gt[0])]}
        index["facts"] = facts
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(GOLD, "workflow_outputs.json"), "w") as fh:
        json.dump(index, fh, indent=1, sort_keys=True)
    print("workflow cases:", sorted(index))


def _jsonable(x):
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating, float)):
        return None if np.isnan(x) else float(x)
    return x


def gen_pairsnp_cases(ref):
    """pairwiseScore (snpmatch.py:270-309) of the reference on npz inputs, with and without a database."""
    rng = np.random.default_rng(21)
    panel = dict(np.load(os.path.join(GOLD, "small_panel.npz")))
    gts = np.array(["0/0", "1/1", "0/1", "1|1", "0|0", "1/0"])
    tmp = tempfile.mkdtemp(prefix="snpm_pairsnp_")
    out, index = {}, {}

    def sample(n_db, n_extra, prefix):
        rows = np.sort(rng.choice(len(panel["positions"]), n_db, replace=False))
        starts = panel["chr_regions"][:, 0]
        chrs = np.char.add(prefix, panel["chrs"].astype("U")[np.searchsorted(starts, rows, side="right") - 1])
        pos = panel["positions"][rows].astype(np.int64)
        # extra markers that are not in the panel, on chromosome 1 and on a contig the panel lacks
        extra_pos = np.setdiff1d(rng.choice(200000, n_extra) + 1, panel["positions"][:int(panel["chr_regions"][0, 1])])
        chrs = np.concatenate([np.repeat(prefix + "1", len(extra_pos)), chrs, np.repeat(prefix + "M", 5)])
        pos = np.concatenate([extra_pos, pos, np.arange(1, 6)])
        order = np.lexsort((pos, chrs))
        chrs, pos = chrs[order], pos[order]
        keep = np.ones(len(pos), dtype=bool)
        keep[1:] = ~((chrs[1:] == chrs[:-1]) & (pos[1:] == pos[:-1]))
        chrs, pos = chrs[keep], pos[keep]
        return chrs, pos, gts[rng.integers(0, 4, len(pos))]

    class _G(ref.snp_genotype.Genotype):
        def __init__(self, hdf5_file, hdf5_acc_file=None):
            g = rh.make_reference_genotype(ref, panel["snps"], panel["positions"], panel["chrs"], panel["chr_regions"], panel["accessions"])
            self.__dict__.update(g.__dict__)

    cases = [(1500, 200, "Chr", 1200, 150, ""), (800, 50, "", 2500, 0, "chr"), (30, 5, "", 40, 5, "")]
    for i, (a, b, pa, c, d, pb) in enumerate(cases):
        c1, p1, g1 = sample(a, b, pa)
        c2, p2, g2 = sample(c, d, pb)
        if i == 2:                                   # disjoint chromosomes on purpose: no common chromosome
            c2 = np.repeat("7", len(c2))
            p2 = np.arange(1, len(c2) + 1)
        f1, f2 = os.path.join(tmp, "s%d_a.npz" % i), os.path.join(tmp, "s%d_b.npz" % i)
        for f, cc, pp, gg in ((f1, c1, p1, g1), (f2, c2, p2, g2)):
            np.savez(f, chr=cc, pos=pp, gt=gg, wei=np.ones((len(pp), 3)), dp=np.ones(len(pp)))
        for k, v in (("c1", c1), ("p1", p1), ("g1", g1), ("c2", c2), ("p2", p2), ("g2", g2)):
            out["p%d_%s" % (i, k)] = v
        index["p%d_plain" % i] = _jsonable(ref.snpmatch.pairwiseScore(f1, f2, False, outFile=None, hdf5File=None))
        orig = ref.snp_genotype.Genotype
        ref.snp_genotype.Genotype = _G
        try:
            r = ref.snpmatch.pairwiseScore(f1, f2, False, outFile=None, hdf5File="panel")
        finally:
            ref.snp_genotype.Genotype = orig
        r.pop("hdf5")
        index["p%d_db" % i] = _jsonable(r)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, "pairsnp.npz"), **out)
    with open(os.path.join(GOLD, "pairsnp.json"), "w") as fh:
        json.dump(index, fh, indent=1, sort_keys=True)
    shutil.rmtree(tmp, ignore_errors=True)
    print("pairsnp: %d cases" % len(cases))


def gen_simulate_cases(ref):
    """simulateSNPs / simulateSNPs_F1 (simulate.py:10-60) of the reference under fixed np.random seeds.  The stand-in
    database carries its accession ids as text: the reference compares them with `str` arguments (simulate.py:12,34), which
    cannot match the bytes array of HDF5Genotype under Python 3."""
    from snpmatch.core import simulate as r_sim
    import pandas as pd
    pd.set_option("future.infer_string", False)     # simulate.py:26 writes integers into a column of strings (object dtype)
    panel = dict(np.load(os.path.join(GOLD, "small_panel.npz")))
    G = rh.make_reference_genotype(ref, panel["snps"], panel["positions"], panel["chrs"], panel["chr_regions"], panel["accessions"])
    G.g.accessions = G.g.accessions.astype("U")
    ids = G.g.accessions
    out = {}
    cases = [("inbred", ids[7], 500, 0.01, 1, 11), ("inbred", ids[0], 64, 0.0, 1, 12), ("inbred", ids[3], 1200, 0.1, 1, 13),
             ("f1", "%sx%s" % (ids[3], ids[21]), 700, 0.01, 1, 14), ("f1", "%sx%s" % (ids[5], ids[6]), 300, 0.05, 0.3, 15)]
    for i, (kind, acc, n, err, rm, seed) in enumerate(cases):
        np.random.seed(seed)
        if kind == "inbred":
            df = r_sim.simulateSNPs(G, str(acc), n, outFile=None, err_rate=err)
            gt = np.array([g.decode() if isinstance(g, bytes) else str(g) for g in df.iloc[:, 2]])
        else:
            # simulate.py:57 assigns the bytes GT strings into an int64 column, which pandas >= 3 refuses.  The conversion
            # is routed around that one assignment: the reference's converter is applied to the returned binary column.
            real = ref.parsers.snp_binary_to_gt
            ref.parsers.snp_binary_to_gt = lambda b: np.array(b)
            try:
                df = r_sim.simulateSNPs_F1(G, str(acc), n, None, err, rm)
            finally:
                ref.parsers.snp_binary_to_gt = real
            gt = np.array([g.decode() for g in real(np.array(df.iloc[:, 2]))])
        out["s%d_kind" % i] = np.array(kind)
        out["s%d_acc" % i] = np.array(str(acc))
        out["s%d_args" % i] = np.array([n, err, rm, seed], dtype=np.float64)
        out["s%d_chr" % i] = np.array(df.iloc[:, 0]).astype("U")
        out["s%d_pos" % i] = np.array(df.iloc[:, 1]).astype(np.int64)
        out["s%d_gt" % i] = gt
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, "simulate.npz"), **out)
    print("simulate: %d cases" % len(cases))


def make_f2_population(panel, parents=(3, 21), n_samples=12, n_db=5200, n_extra=150, seed=31):
    """A multi-sample VCF worth of arrays: F2-like mosaics (parent 1 / het / parent 2 in 2-4 Mb blocks) of two panel accessions
    on a subset of the panel's positions plus positions the panel lacks, with call errors and missing calls."""
    rng = np.random.default_rng(seed)
    rows = np.sort(rng.choice(len(panel["positions"]), n_db, replace=False))
    starts = panel["chr_regions"][:, 0]
    chr_of = np.searchsorted(starts, rows, side="right") - 1
    chrs = np.char.add("Chr", panel["chrs"].astype("U")[chr_of])
    pos = panel["positions"][rows].astype(np.int64)
    p1, p2 = panel["snps"][rows, parents[0]], panel["snps"][rows, parents[1]]
    gt = np.zeros((n_db, n_samples), dtype=np.int8)
    for s in range(n_samples):
        block = (pos // int(rng.integers(2_000_000, 4_000_000))) + 7 * chr_of + s
        state = (block * 2654435761 % 4)                      # 0: parent 1, 1/2: het, 3: parent 2
        het_ok = (p1 >= 0) & (p2 >= 0) & (p1 != p2)
        call = np.where(state == 0, p1, np.where(state == 3, p2, np.where(het_ok, 2, p1)))
        call = np.where(call < 0, 0, call)
        err = rng.random(n_db) < 0.01
        call = np.where(err, rng.integers(0, 3, n_db), call)
        miss = rng.random(n_db) < (0.05 if s != n_samples - 1 else 0.97)   # the last sample is almost empty
        gt[:, s] = np.where(miss, -1, call)
    # positions the panel lacks, chromosome 1, and a contig outside the genome would trip the reference's assert: none added
    extra = np.setdiff1d(rng.choice(30_000_000, n_extra) + 1, panel["positions"][:int(panel["chr_regions"][0, 1])])
    chrs = np.concatenate([np.repeat("Chr1", len(extra)), chrs])
    pos = np.concatenate([extra, pos])
    gt = np.concatenate([rng.integers(-1, 3, size=(len(extra), n_samples)).astype(np.int8), gt])
    order = np.lexsort((pos, chrs))
    names = np.array(["./.", "0/0", "1/1", "0/1"])
    return {"samples": np.array(["F2_%02d" % s for s in range(n_samples)]), "chr": chrs[order], "pos": pos[order],
            "gt": names[gt[order].astype(int) + 1]}


def gen_genotype_cross_cases(ref):
    """GenotypeCross.genotype_cross (genotype_cross.py:210-241) of the reference; the VCF parse (scikit-allel, absent here) is
    replaced by arrays of the shape import_vcf_file returns (parsers.py:176-213)."""
    from snpmatch.core import genotype_cross as r_gc
    panel = dict(np.load(os.path.join(GOLD, "small_panel.npz")))
    G = rh.make_reference_genotype(ref, panel["snps"], panel["positions"], panel["chrs"], panel["chr_regions"], panel["accessions"])
    ids = G.accessions
    vcf = make_f2_population(panel)
    np.savez_compressed(os.path.join(GOLD, "genotype_cross_vcf.npz"), **vcf)
    r_gc.genome = ref.genomes.Genome("athaliana_tair10")
    real = ref.parsers.import_vcf_file
    ref.parsers.import_vcf_file = lambda inFile, logDebug=False, samples_to_load=None, add_fields=None: dict(vcf)
    index = {}
    try:
        for tag, bin_len, lr in (("b300k_lr1.5", 300000, 1.5), ("b1M_lr3", 1000000, 3.0), ("b2M_lr1.5", 2000000, 1.5)):
            gc = r_gc.GenotypeCross(G, "%sx%s" % (ids[3], ids[21]), bin_len, None, False)
            lines = gc.genotype_cross("population.vcf", lr)
            index[tag] = {"bin_len": bin_len, "lr_thres": lr, "parents": "%sx%s" % (ids[3], ids[21]),
                          "n_segregating": int(len(gc.commonSNPsPOS)), "lines": [str(x) for x in lines]}
    finally:
        ref.parsers.import_vcf_file = real
    # known answers of getWindowGenotype on a grid of counts
    grid = []
    for total in (3, 5, 8, 20, 57):
        for a in range(0, total + 1, max(1, total // 6)):
            for h in range(0, total + 1 - a, max(1, total // 5)):
                for b in (0, total - a - h, (total - a - h) // 2):
                    for lr in (1.5, 3.0):
                        geno, pval = r_gc.getWindowGenotype([a, h, b], total, lr)
                        grid.append([total, a, h, b, lr, -1 if geno == "NA" else int(geno)])
    index["window_genotype_grid"] = grid
    with open(os.path.join(GOLD, "genotype_cross.json"), "w") as fh:
        json.dump(index, fh, indent=0, sort_keys=True)
    print("genotype_cross: %d cases, grid of %d cells" % (len(index) - 1, len(grid)))


def gen_makedb_cases(ref):
    """The CSV loader behind `snpmatch makedb` (pygwas/genotype.py:29-105: parse_genotype_csv_file, load_csv_genotype_data) on a
    small CSV: chromosome entries per run of labels in file order (a label that comes back opens a new entry)."""
    from snpmatch.pygwas import genotype as r_pg
    rng = np.random.default_rng(41)
    labels = ["1"] * 7 + ["2"] * 4 + ["Chr5"] * 5 + ["2"] * 3 + ["M"]
    pos = np.concatenate([np.sort(rng.choice(5000, n, replace=False)) + 1 for n in (7, 4, 5, 3, 1)])
    codes = rng.choice(np.array([-1, 0, 1, 2]), size=(len(labels), 6), p=[0.1, 0.5, 0.3, 0.1])
    accs = ["6909", "8236", " 7000", "9001", "100", "5"]
    text = "Chromosome,Position," + ",".join(accs) + "\n" + "".join(
        "%s,%d,%s\n" % (c, p, ",".join(str(v) for v in row)) for c, p, row in zip(labels, pos, codes))
    tmp = tempfile.mkdtemp(prefix="snpm_makedb_")
    path = os.path.join(tmp, "db.csv")
    with open(path, "w") as fh:
        fh.write(text)
    g = r_pg.load_csv_genotype_data(path)
    out = {"csv": np.array(text), "snps": np.array(g.snps, dtype=np.int8), "positions": np.array(g.positions, dtype=np.int64),
           "chrs": np.array(g.chrs, dtype="U"), "chr_regions": np.array(g.chr_regions, dtype=np.int64),
           "accessions": np.array(g.accessions, dtype="U")}
    np.savez_compressed(os.path.join(GOLD, "makedb_csv.npz"), **out)
    shutil.rmtree(tmp, ignore_errors=True)
    print("makedb: %d rows, %d chromosome entries" % (len(out["positions"]), len(out["chrs"])))


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = rh.load_reference()
    gen_match_cases(ref)
    gen_join_cases(ref)
    gen_epilogue_cases(ref)
    gen_workflow_cases(ref)
    gen_pairsnp_cases(ref)
    gen_simulate_cases(ref)
    gen_genotype_cross_cases(ref)
    gen_makedb_cases(ref)


if __name__ == "__main__":
    main()
