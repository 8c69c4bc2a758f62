"""
`snpmatch inbred` — host side of the B200 matching path.

Mirrors the public surface of the reference's `snpmatch/core/snpmatch.py`: module constants
(:17-19), `get_fraction` (:25-28), `likeliTest` (:40-55), `np_test_identity` (:57-72),
`matchGTsAccs` (:74-89), `GenotyperOutput` (:91-168), `Genotyper` (:170-241), `getHeterozygosity`
(:244-253), `potatoGenotyper` (:256-268), `pairwiseScore` (:270-309).  All array arithmetic of the path — the join, the
per-accession reduction, truncation, probabilities, likelihoods and ratios — runs in
libsnpmatch_b200 on the GPU (snpmatch_b200/csrc); this file only prepares buffers and writes the
reference's output files (`scores.txt`, `matches.json`).
"""
import json
import logging
import os
import sys

import numpy as np
import pandas as pd

from .. import lib
from . import labels
from . import parsers
from . import snp_genotype

log = logging.getLogger(__name__)
lr_thres = 3.841
snp_thres = 4000
prob_thres = 0.98


def die(msg):
    sys.stderr.write('Error: ' + msg + '\n')
    sys.exit(1)


def get_fraction(x, y, y_min=0):
    if y <= y_min:
        return np.nan
    return float(x) / y


np_get_fraction = np.vectorize(get_fraction, excluded="y_min")


def likeliTest(n, y):
    """Likelihood of y matches out of n informative sites (snpmatch.py:40-55), evaluated by the
    device epilogue (csrc/score.cuh: likeli_test) so that scalars and arrays agree bit for bit."""
    assert y <= n, "provided y is greater than n"
    _, lik, _ = lib.calculate_likelihoods(np.array([y], dtype=np.float64), np.array([n], dtype=np.float64))
    return float(lik[0])


_KMAX_CACHE = {}


def identity_kmax_table(n_max, error_rate=0.02, pthres=0.05):
    """kmax[n] = largest k with binom.sf(k - 1, n, error_rate) >= pthres, for n = 0..n_max.

    np_test_identity (snpmatch.py:57-72) is `binom.sf((n - x) - 1, n, e) >= pthres`; SciPy floors
    the first argument and sf is non-increasing in it, so identical <=> floor(n - x - 1) + 1 <=
    kmax[n].  The table is built with the same SciPy function the reference calls, which makes the
    device's identity calls bit-exact by construction."""
    from scipy import stats
    key = (float(error_rate), float(pthres))
    have = _KMAX_CACHE.get(key)
    if have is not None and len(have) > n_max:
        return have
    n_max = max(int(n_max), 64)
    n = np.arange(n_max + 1)
    # candidate from the inverse survival function, then fix up with sf itself
    k = np.clip(stats.binom.isf(pthres, n, error_rate).astype(np.int64) + 1, 0, n + 1)
    for _ in range(64):
        ok_here = stats.binom.sf(k - 1, n, error_rate) >= pthres
        ok_next = stats.binom.sf(k, n, error_rate) >= pthres
        up = ok_here & ok_next & (k < n + 1)
        down = ~ok_here & (k > 0)
        if not (up.any() or down.any()):
            break
        k = np.where(up, k + 1, np.where(down, k - 1, k))
    table = k.astype(np.int32)
    _KMAX_CACHE[key] = table
    return table


def np_test_identity(x, n, error_rate=0.0005, pthres=0.05):
    """snpmatch.py:70-72 through the kmax table (host integers; the windowed path evaluates the same
    comparison on the device)."""
    x = np.asarray(x, dtype=np.float64)
    n = np.asarray(n)
    table = identity_kmax_table(int(np.max(n)) if n.size else 0, error_rate, pthres)
    return np.array(np.floor(n - x - 1) + 1 <= table[n.astype(np.int64)]).astype(int)


def matchGTsAccs(sampleWei, t1001snps, skip_hets_db=False):
    """Weighted genotype match of k markers against every accession (snpmatch.py:74-89) on the GPU:
    returns (score f64[A], ninfo int[A]) in the reference's summation order."""
    sampleWei = np.asarray(sampleWei)
    t1001snps = np.asarray(t1001snps)
    assert sampleWei.shape[0] == t1001snps.shape[0], "please provide same number of positions for both sample and db"
    assert sampleWei.ndim == 2 and sampleWei.shape[1] == 3, "SNP weights should be a np.array with  shape == n,3"
    return lib.match_gts_accs(sampleWei, t1001snps, skip_hets_db)


class GenotyperOutput(object):
    """Result table of a run (snpmatch.py:91-168)."""

    def __init__(self, AccList, ScoreList, NumInfoSites, overlap, NumMatSNPs, DPmean):
        self._fused = None
        self.overlap, self.num_snps, self.dp = overlap, NumMatSNPs, DPmean
        self.accs = np.asarray(AccList).astype("str")
        self.ninfo = np.asarray(NumInfoSites).astype("int")
        self.scores = np.asarray(ScoreList).astype("int")     # truncation toward zero (snpmatch.py:96)

    def _attach_fused(self, prob, lik, lrt):
        """Results of the epilogue that ran fused behind the scoring kernels."""
        self._fused = (len(self.accs), np.array(prob), np.array(lik), np.array(lrt))

    def _fused_ok(self):
        return self._fused is not None and self._fused[0] == len(self.scores)

    def get_probabilities(self):
        if self._fused_ok():
            self.probabilies = self._fused[1]
        else:
            self.probabilies, _, _ = lib.calculate_likelihoods(self.scores, self.ninfo)

    @staticmethod
    def calculate_likelihoods(scores, ninfo, amin="calc"):
        """(L, LR) of snpmatch.py:106-117, computed by the device epilogue."""
        _, lik, lrt = lib.calculate_likelihoods(scores, ninfo, amin)
        return lik, lrt

    def get_likelihoods(self, amin="calc"):
        if amin == "calc" and self._fused_ok():
            self.likelis, self.lrts = self._fused[2], self._fused[3]
        else:
            self.likelis, self.lrts = self.calculate_likelihoods(self.scores, self.ninfo, amin)

    def print_out_table(self, outFile):
        """`scores.txt`: acc, matches, ninfo, probability, likelihood, lrt, num_snps, dp; tab-separated,
        no header (snpmatch.py:122-138).  BED inputs carry dp = "NA": the mean is nan."""
        self.get_likelihoods()
        self.get_probabilities()
        columns = (('accs', self.accs), ('matches', self.scores), ('ninfo', self.ninfo), ('probabilities', self.probabilies),
                   ('likelihood', self.likelis), ('lrt', self.lrts), ('num_snps', self.num_snps), ('dp', parsers.mean_depth(self.dp)))
        n_rows = len(self.accs)
        table = pd.DataFrame({name: (col if np.ndim(col) else np.repeat(col, n_rows)) for name, col in columns},
                             columns=[name for name, _ in columns])
        if outFile:
            table.to_csv(outFile, sep="\t", header=False, index=False)
        return table

    def print_json_output(self, outFile):
        """`matches.json` (snpmatch.py:140-150)."""
        self.get_likelihoods()
        self.get_probabilities()
        with np.errstate(invalid="ignore"):
            topHits = np.where(self.lrts < lr_thres)[0]
        by_probability = topHits[np.argsort(-self.probabilies[topHits])]
        case, note = self.case_interpreter(topHits)
        rows = []
        for i in by_probability:            # accession, probability, informative sites, informative sites / matched markers
            rows.append((str(self.accs[i]), float(self.probabilies[i]), int(self.ninfo[i]), float(get_fraction(self.ninfo[i], self.num_snps))))
        report = {'interpretation': {'case': case, 'text': note}, 'matches': rows, 'overlap': [self.overlap, self.num_snps]}
        with open(outFile, "w") as fh:
            json.dump(report, fh, sort_keys=True, indent=4)

    def case_interpreter(self, topHits):
        """snpmatch.py:152-168."""
        overlap_thres = 0.5
        case, note = 1, "Ambiguous sample"
        if len(topHits) == 1:
            case, note = 0, "Unique hit"
        elif np.nanmean(self.probabilies[topHits]) > prob_thres:
            case, note = 2, "Ambiguous sample: Accessions in top hits can be really close"
        elif self.overlap > overlap_thres:
            case, note = 3, "Ambiguous sample: Sample might contain mixture of DNA or contamination"
        elif self.overlap < overlap_thres:
            case, note = 4, "Ambiguous sample: Many input SNP positions are missing in db positions. Maybe sample  not one in database"
        return case, note


class Genotyper(object):
    """`snpmatch inbred` for one sample against the resident database (snpmatch.py:170-241)."""

    def __init__(self, inputs, g, outFile, run_genotyper=True, skip_db_hets=False, chunk_size=1000):
        assert type(g) is snp_genotype.Genotype, "provide a snp_genotype.Genotype class for genotypes"
        assert int(chunk_size) >= 1, "chunk_size must be a positive number of SNPs"
        self.g, self.inputs, self.outFile = g, inputs, outFile
        self.chunk_size, self._skip_db_hets = chunk_size, skip_db_hets
        self.num_lines = len(g.g.accessions)
        self.inputs.filter_chr_names()
        if run_genotyper:
            self.result = self.genotyper()
            self.write_genotyper_output(self.result)

    def get_common_positions(self):
        self.commonSNPs = self.g.get_positions_idxs(self.inputs.chrs, self.inputs.pos)

    def filter_tophits(self):
        """`--refine` (snpmatch.py:189-205): rescoring over the SNPs that segregate among the
        indistinguishable accessions."""
        self.result = self.genotyper()
        self.write_genotyper_output(self.result)
        self.result.get_likelihoods()
        with np.errstate(invalid="ignore"):
            topHits = np.where(self.result.lrts < lr_thres)[0]
        if len(topHits) == 1:
            log.info("a single accession is left: perfect hit, nothing to refine")
            return None
        log.info("%s accessions cannot be told apart at LR < %s", len(topHits), lr_thres)
        if len(topHits) > (self.num_lines / 2):
            log.info("more than half of the panel is indistinguishable: not refining")
            return None
        seg_ix = identify_segregating_snps(self.g, topHits)
        with np.errstate(invalid="ignore"):
            mask = np.where(self.result.lrts >= lr_thres)[0]
        self.result_fine = self.genotyper(filter_pos_ix=seg_ix, mask_acc_ix=mask)
        self.result_fine.print_out_table(self.outFile + ".refined.scores.txt")

    def genotyper(self, filter_pos_ix=None, mask_acc_ix=None):
        """Join + chunked scoring + epilogue in one device pass (snpmatch.py:207-233)."""
        g = self.g
        order, cid, pos = g.prepare_markers(self.inputs.chrs, self.inputs.pos)
        wei = np.ascontiguousarray(np.asarray(self.inputs.wei, dtype=np.float64)[order])
        # called genotypes (BED, VCF without PL) take the popcount kernel; likelihood-weighted samples the fp64 one, which
        # sums in the reference's chunks of `chunk_size` rows (snpmatch.py:173,218: the chunking is part of its rounding)
        mode = lib.KERNEL_POPCOUNT if lib.weights_are_one_hot(wei) else lib.KERNEL_FP64
        batch = g.db.scratch_batch([0, len(pos)], cid, pos, wei,
                                   chunk_rows=int(self.chunk_size) if mode == lib.KERNEL_FP64 else lib.CHUNK_ROWS)
        try:
            if filter_pos_ix is not None:
                assert type(filter_pos_ix) is np.ndarray, "provide np array for indices to be considered"
                batch.set_row_filter(filter_pos_ix)
            batch.run(self._skip_db_hets, kernel_mode=mode)
            batch.epilogue()
            r = batch.fetch()
            db_idx, s_idx = batch.fetch_pairs(0)
            self.timings = batch.timings()
        finally:
            batch.close()
        if filter_pos_ix is not None and len(db_idx) < 100:
            log.info("#positions in segregating sites are are too little: %s" % len(db_idx))
        self.commonSNPs = (db_idx, order[s_idx])
        NumMatSNPs = int(r["m"][0])
        overlap = get_fraction(NumMatSNPs, len(self.inputs.pos))
        accs = g.g.accessions
        if mask_acc_ix is not None:
            assert type(mask_acc_ix) is np.ndarray, "provide a numpy array of accessions indices to mask"
            keep = np.setdiff1d(np.arange(self.num_lines), mask_acc_ix)
            return GenotyperOutput(accs[keep], r["score"][0][keep], r["ninfo"][0][keep], overlap, NumMatSNPs, self.inputs.dp)
        out = GenotyperOutput(accs, r["score"][0], r["ninfo"][0], overlap, NumMatSNPs, self.inputs.dp)
        out._attach_fused(r["prob"][0], r["L"][0], r["LR"][0])
        return out

    def write_genotyper_output(self, result):
        log.info("writing %s.scores.txt and %s.matches.json", self.outFile, self.outFile)
        result.get_likelihoods()
        result.print_out_table(self.outFile + '.scores.txt')
        result.print_json_output(self.outFile + ".matches.json")
        getHeterozygosity(self.inputs.gt[self.commonSNPs[1]], self.outFile + ".matches.json")
        return result


def identify_segregating_snps(g, accs_ix):
    """Rows where the given accessions carry more than one distinct called genotype (snp_genotype.py:188-211 with
    segregting_snps :378-383).  One scan of the resident panel on the GPU.  Like the reference, returns None when more
    than half of the accessions are selected."""
    assert type(accs_ix) is np.ndarray, "provide an np array for list of indices to be considered"
    assert len(accs_ix) > 1, "polymorphism happens in more than 1 line"
    if len(accs_ix) > (len(g.accessions) / 2):
        return None
    return g.db.segregating_rows(accs_ix) + g.db.row0_global


def getHeterozygosity(snpGT, outFile='default'):
    """Fraction of heterozygous calls among the matched markers; added to the JSON (snpmatch.py:244-253)."""
    snpBinary = parsers.parseGT(snpGT)
    numHets = int(np.count_nonzero(snpBinary == 2))
    frac = get_fraction(numHets, len(snpGT))
    if outFile != 'default':
        with open(outFile) as fh:
            report = json.load(fh)
        report['percent_heterozygosity'] = frac
        with open(outFile, "w") as fh:
            json.dump(report, fh, sort_keys=True, indent=4)
    return frac


def potatoGenotyper(args):
    inputs = parsers.ParseInputs(inFile=args['inFile'], logDebug=args['logDebug'])
    log.info("packing the database into HBM")
    g = snp_genotype.Genotype(args['hdf5File'], args['hdf5accFile'])
    log.info("matching the sample against %s accessions", len(g.accessions))
    genotyper = Genotyper(inputs, g, args['outFile'], run_genotyper=not args['refine'], skip_db_hets=args['skip_db_hets'])
    if args['refine']:                       # --refine: score, then re-score the indistinguishable accessions on their segregating SNPs
        genotyper.filter_tophits()
    log.info("finished!")


def pairwiseScore(inFile_1, inFile_2, logDebug, outFile=None, hdf5File=None, device=0):
    """`snpmatch pairsnp` (snpmatch.py:270-309): fraction of identical genotype strings among the markers two samples
    share, per chromosome and overall, optionally restricted to the positions of a database.  Both joins and the counting
    loop (snpmatch.py:291-297) run on the GPU; the returned dict has the reference's keys and values.  Unlike the reference
    under Python 3 (its json.dumps trips over numpy integers, snpmatch.py:307) the output file is written."""
    snpmatch_stats = {}
    log.info("parsing %s and %s", inFile_1, inFile_2)
    inputs_1 = parsers.ParseInputs(inFile=inFile_1, logDebug=logDebug)
    inputs_2 = parsers.ParseInputs(inFile=inFile_2, logDebug=logDebug)
    if hdf5File is not None:
        log.info("restricting sample 1 to the positions of the database")
        g = hdf5File if isinstance(hdf5File, snp_genotype.Genotype) else snp_genotype.Genotype(hdf5File, None, device=device)
        snpmatch_stats['hdf5'] = hdf5File if not isinstance(hdf5File, snp_genotype.Genotype) else "resident"
        commonSNPs_1 = g.get_positions_idxs(inputs_1.chrs, inputs_1.pos)
        common_inds = snp_genotype.Genotype.get_common_positions(inputs_1.chrs[commonSNPs_1[1]], inputs_1.pos[commonSNPs_1[1]],
                                                                 inputs_2.chrs, inputs_2.pos, device=device)
        common_inds = (commonSNPs_1[1][common_inds[0]], common_inds[1])
    else:
        log.info("joining the two samples on (chromosome, position)")
        common_inds = snp_genotype.Genotype.get_common_positions(inputs_1.chrs, inputs_1.pos, inputs_2.chrs, inputs_2.pos, device=device)
    log.info("done!")
    n1, n2 = len(inputs_1.chrs), len(inputs_2.chrs)
    unique_1 = n1 - len(common_inds[0])
    unique_2 = n2 - len(common_inds[0])
    inputs_1.filter_chr_names()
    inputs_2.filter_chr_names()
    common_chrs = np.intersect1d(inputs_1.g_chrs_ids, inputs_2.g_chrs_ids)
    # what crosses the C ABI: chromosome id (index into common_chrs, -1 elsewhere) of every marker of sample 1 and the ids
    # of the genotype strings in one table for both samples
    codes1, uniq1 = labels.factorize(inputs_1.g_chrs)
    in_common = {c: k for k, c in enumerate(common_chrs)}
    chrom1 = np.array([in_common.get(u, -1) for u in uniq1], dtype=np.int32)[codes1] if n1 else np.zeros(0, dtype=np.int32)
    gt_ids = labels.factorize(np.concatenate([inputs_1.gt.astype("U"), inputs_2.gt.astype("U")]))[0]
    common, scores = lib.pair_match_counts(common_inds[0], common_inds[1], chrom1, gt_ids[:n1], gt_ids[n1:], len(common_chrs), device=device)
    for k, i in enumerate(common_chrs):
        log.debug("chromosome %s: %s common markers, %s identical calls", i, int(common[k]), int(scores[k]))
        snpmatch_stats[str(i)] = [get_fraction(int(scores[k]), int(common[k])), int(common[k])]
    snpmatch_stats['matches'] = [get_fraction(int(np.sum(scores)), int(np.sum(common))), int(np.sum(common))]
    snpmatch_stats['unique'] = {"%s" % os.path.basename(inFile_1): [get_fraction(unique_1, n1), n1],
                                "%s" % os.path.basename(inFile_2): [get_fraction(unique_2, n2), n2]}
    if outFile:
        log.info("writing %s.matches.json", outFile)
        with open(outFile + ".matches.json", "w") as fh:
            json.dump(snpmatch_stats, fh, sort_keys=True, indent=4)
    return snpmatch_stats
