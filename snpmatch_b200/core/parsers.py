"""
Sample parsing and PL -> weight preparation (host side; stays Python per the north star).

Mirrors the interface of the reference's `snpmatch/core/parsers.py` — `ParseInputs`
(parsers.py:59-175), `parseGT` (:12-35), `snp_binary_to_gt` (:37-44), `import_vcf_file`
(:178-213), `potatoParser` (:216-218) — and produces the buffers that cross the C ABI
(`chrs`, `pos`, `gt`, `wei` f64 [n,3], `dp`).

The reference delegates VCF reading to scikit-allel (parsers.py:184,189), which is not
available here; `read_vcf_minimal` is an in-repo reader that reproduces the fields the path
uses (first sample's GT and PL, CHROM, POS, INFO/DP).  It is validated against the golden
facts of the reference's sample files.

Documented deviation: BED inputs carry dp = "NA"; the reference crashes
in np.nanmean on the first parse (parsers.py:113) — here the depth statistic is nan instead.
"""
import gzip
import json
import logging
import os
import re
import sys

import numpy as np

log = logging.getLogger(__name__)

snp_thres = 4000   # snpmatch.py:18


def die(msg):
    sys.stderr.write('Error: ' + msg + '\n')
    sys.exit(1)


def parseGT(snpGT):
    """GT strings -> int8 codes: 1/1 -> 1, 0/1|1/0 -> 2, ./. -> -1, anything else 0.
    The separator is sniffed from the first element only (parsers.py:12-35)."""
    snpGT = np.asarray(snpGT)
    codes = np.zeros(len(snpGT), dtype="int8")
    if len(codes) == 0:
        return codes
    first = str(snpGT[0])
    if "|" in first:
        sep = "|"
    elif "/" in first:
        sep = "/"
    elif first.isdigit():
        return np.array(snpGT, dtype="int8")
    else:
        die("unable to parse the format of GT in vcf!")
    gt = snpGT.astype("str")
    codes[gt == "1" + sep + "1"] = 1
    codes[(gt == "0" + sep + "1") | (gt == "1" + sep + "0")] = 2
    codes[gt == "." + sep + "."] = -1
    return codes


def snp_binary_to_gt(snpBinary):
    """int8 codes -> GT byte strings (parsers.py:37-44)."""
    snpBinary = np.array(snpBinary, dtype="int8")
    table = {-1: b"./.", 0: b"0/0", 1: b"1/1", 2: b"0/1"}
    out = np.zeros(len(snpBinary), dtype="S8")
    for code, text in table.items():
        out[snpBinary == code] = text
    return out


def mean_depth(dp):
    """np.nanmean(dp) that yields nan instead of raising for the "NA" depth of BED inputs."""
    try:
        arr = np.asarray(dp, dtype=float)
    except (TypeError, ValueError):
        return float("nan")
    if arr.size == 0 or np.all(np.isnan(arr)):
        return float("nan")
    return float(np.nanmean(arr))


def read_vcf_minimal(inFile, sample_index=0):
    """Minimal VCF reader: CHROM, POS, INFO/DP, and GT + PL of one sample.

    Output conventions follow what the reference receives from scikit-allel
    (parsers.py:191-206): GT rendered as 'a/b' with '.' for a missing allele (phasing is
    dropped by GenotypeArray.to_gt), PL as float [n,3] with -1 for missing entries,
    dp = INFO/DP (-1 when absent on a record) or an array of "NA" when the header declares
    no INFO DP field."""
    opener = gzip.open if str(inFile).endswith(".gz") else open
    chrom, pos, gts, pls, dps = [], [], [], [], []
    has_info_dp = False
    has_pl = False
    dp_re = re.compile(r"(?:^|;)DP=([^;]+)")
    with opener(inFile, "rt") as fh:
        for line in fh:
            if line.startswith("##"):
                if line.startswith("##INFO=<ID=DP,"):
                    has_info_dp = True
                continue
            if line.startswith("#"):
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 10 + sample_index:
                continue
            chrom.append(f[0])
            pos.append(int(f[1]))
            m = dp_re.search(f[7])
            try:
                dps.append(int(m.group(1)) if m else -1)
            except ValueError:
                dps.append(-1)
            keys = f[8].split(":")
            vals = f[9 + sample_index].split(":")
            rec = dict(zip(keys, vals))
            alleles = re.split(r"[/|]", rec.get("GT", "."))
            if len(alleles) == 1:
                alleles = [alleles[0], "."] if alleles[0] != "." else [".", "."]
            gts.append("/".join(a if a != "" else "." for a in alleles[:2]))
            pl = [-1.0, -1.0, -1.0]
            if "PL" in rec:
                has_pl = True
                for i, v in enumerate(rec["PL"].split(",")[:3]):
                    if v not in (".", ""):
                        pl[i] = float(v)
            elif "PL" in keys:
                has_pl = True
            pls.append(pl)
    out = {
        "chr": np.array(chrom, dtype="str"),
        "pos": np.array(pos, dtype=np.int64),
        "gt": np.array(gts, dtype="str"),
    }
    if has_pl:
        out["wei"] = np.array(pls, dtype=float).reshape(-1, 3)
    out["dp"] = np.array(dps, dtype=np.int64) if has_info_dp else np.repeat("NA", len(pos))
    return out


def read_vcf_samples(inFile):
    """All samples of a VCF: 'samples' str[S], 'chr', 'pos', 'gt' str [n,S] rendered 'a/b' as scikit-allel's
    GenotypeArray.to_gt does for the reference (parsers.py:191-193).  Used by genotype_cross (genotype_cross.py:212)."""
    opener = gzip.open if str(inFile).endswith(".gz") else open
    samples, chrom, pos, rows = [], [], [], []
    with opener(inFile, "rt") as fh:
        for line in fh:
            if line.startswith("##"):
                continue
            f = line.rstrip("\n").split("\t")
            if line.startswith("#"):
                samples = f[9:]
                continue
            if len(f) < 10:
                continue
            chrom.append(f[0])
            pos.append(int(f[1]))
            gt_at = f[8].split(":").index("GT") if "GT" in f[8].split(":") else -1
            row = []
            for cell in f[9:9 + len(samples)]:
                v = cell.split(":")
                raw = v[gt_at] if 0 <= gt_at < len(v) else "."
                alleles = re.split(r"[/|]", raw)
                if len(alleles) == 1:
                    alleles = [alleles[0], "."] if alleles[0] != "." else [".", "."]
                row.append("/".join(a if a != "" else "." for a in alleles[:2]))
            row += ["./."] * (len(samples) - len(row))
            rows.append(row)
    return {"samples": np.array(samples, dtype="str"), "chr": np.array(chrom, dtype="str"), "pos": np.array(pos, dtype=np.int64),
            "gt": np.array(rows, dtype="str").reshape(len(pos), len(samples))}


def import_vcf_file(inFile, logDebug=False, samples_to_load=[0], add_fields=None):
    """Same role as parsers.py:178-213; returns dict with 'gt' [n,1], 'wei' [n,1,3], 'chr', 'pos', 'dp'.
    samples_to_load=None loads the genotypes of every sample ('samples', 'gt' [n,S]; genotype_cross.py:212)."""
    if samples_to_load is None:
        out = read_vcf_samples(inFile)
        out["dp"] = np.repeat("NA", len(out["pos"]))
        return out
    raw = read_vcf_minimal(inFile, sample_index=samples_to_load[0])
    snp_inputs = {"chr": raw["chr"], "pos": raw["pos"], "dp": raw["dp"], "gt": raw["gt"][:, None]}
    if "wei" in raw:
        snp_inputs["wei"] = raw["wei"][:, None, :]
    return snp_inputs


class ParseInputs(object):
    """VCF / BED / npz -> chrs, pos, gt, wei, dp (parsers.py:59-175)."""

    def __init__(self, inFile, logDebug=True, outFile="parser"):
        if outFile == "parser" or not outFile:
            outFile = inFile + ".snpmatch"
        if os.path.isfile(inFile + ".snpmatch.npz"):
            log.info("snpmatch parser dump found! loading %s", inFile + ".snpmatch.npz")
            snps = np.load(inFile + ".snpmatch.npz")
            self.load_snp_info(snps['chr'], snps['pos'], snps['gt'], snps['wei'], snps['dp'])
        elif os.path.isfile(inFile):
            _, inType = os.path.splitext(inFile)
            if inType == '.npz':
                snps = np.load(inFile)
                self.load_snp_info(snps['chr'], snps['pos'], snps['gt'], snps['wei'], snps['dp'])
            else:
                if inType == '.vcf' or os.path.basename(inFile).endswith(".vcf.gz"):
                    parsed = self.read_vcf(inFile, logDebug)
                elif inType == '.bed':
                    parsed = self.read_bed(inFile, logDebug)
                else:
                    die("input file type %s not supported" % inType)
                self.load_snp_info(*parsed)
                self.save_snp_info(outFile)
                self.case_interpret_inputs(outFile + ".stats.json")

    def load_snp_info(self, snpCHR, snpPOS, snpGT, snpWEI, DPmean):
        self.chrs = np.array(snpCHR, dtype="str")
        self.pos = np.array(snpPOS, dtype=int)
        self.gt = np.array(snpGT, dtype="str")
        self.wei = np.array(snpWEI, dtype=float)
        self.dp = DPmean
        pc = getattr(self, "_pl_codes", None)       # read_vcf: integer PLs as weight codes
        self._coded = (pc[0], pc[1], self.wei) if pc is not None and len(pc[0]) == len(self.wei) else None
        self._pl_codes = None

    def coded_weights(self):
        """(codes uint16 [n,3], table f64 [V]) with table[codes] == wei bit for bit, or None when the weights take more than
        65536 distinct values.  A VCF with integer PLs gives the codes for free (code = PL, table = exp(-PL/10): read_vcf keeps
        them); any other input is dictionary-coded once here and cached.  This is what crosses the PCIe bus in the batched path
        (lib.CodedSamples): 6 instead of 24 bytes per marker."""
        cached = getattr(self, "_coded", None)
        if cached is None or cached[2] is not self.wei:
            from .. import lib
            iw = lib.index_weights(self.wei)
            cached = (iw[0], iw[1], self.wei) if iw is not None else (None, None, self.wei)
            self._coded = cached
        return None if cached[0] is None else (cached[0], cached[1])

    def save_snp_info(self, outFile):
        np.savez(outFile, chr=self.chrs, pos=self.pos, gt=self.gt, wei=self.wei, dp=self.dp)

    def case_interpret_inputs(self, outFile):
        from . import snpmatch as _sm
        n = len(self.chrs)
        names, counts = np.unique(self.chrs, return_counts=True)
        stat = {
            "snps": dict((str(k), int(v)) for k, v in zip(names, counts)),
            "interpretation": {"case": 1, "text": "Attention: low number of SNPs provided"} if n < snp_thres
            else {"case": 0, "text": "Sufficient number of SNPs"},
            "num_of_snps": n,
            "depth": mean_depth(self.dp),
            "percent_heterozygosity": _sm.getHeterozygosity(self.gt),
        }
        with open(outFile, "w") as out_stats:
            out_stats.write(json.dumps(stat))

    @staticmethod
    def get_wei_from_GT(snpGT):
        """One-hot weights of the called genotype, columns (0/0, 0/1, 1/1) (parsers.py:132-139)."""
        codes = parseGT(snpGT)
        wei = np.zeros((len(codes), 3), dtype=float)
        wei[codes == 0, 0] = 1.0
        wei[codes == 2, 1] = 1.0
        wei[codes == 1, 2] = 1.0
        return wei

    @staticmethod
    def read_bed(inFile, logDebug):
        """Three whitespace-separated columns chr, pos, GT (parsers.py:118-130)."""
        chrs, pos, gt = [], [], []
        with open(inFile) as fh:
            for line in fh:
                f = line.split()
                if len(f) < 3:
                    continue
                chrs.append(f[0])
                pos.append(int(f[1]))
                gt.append(f[2])
        gt = np.array(gt)
        return (np.array(chrs, dtype="str"), np.array(pos, dtype=int), gt,
                ParseInputs.get_wei_from_GT(gt), "NA")

    def read_vcf(self, inFile, logDebug):
        """First sample; drop no-calls; wei = exp(-PL/10), GT one-hot where PL is absent
        (parsers.py:141-157)."""
        v = import_vcf_file(inFile, logDebug, samples_to_load=[0])
        gt_all = v['gt'][:, 0]
        req = np.flatnonzero((gt_all != './.') & (gt_all != '.|.'))
        gt = gt_all[req]
        self._pl_codes = None
        if 'wei' in v:
            pl = v['wei'][req, 0]
            no_pl = np.all(pl == -1, axis=1)
            wei = np.exp(pl / (-10))
            wei[no_pl, ] = self.get_wei_from_GT(gt[no_pl])
            # integer PLs are ready-made dictionary codes of the weights: table[k] = exp(k / -10), one extra entry for 0.0
            if len(pl) and np.all(pl == np.floor(pl)) and pl.min() >= -1 and pl.max() < 65000:
                top = int(pl.max()) + 1
                table = np.append(np.exp(np.arange(top) / (-10)), 0.0)
                codes = np.where(pl < 0, 0, pl).astype(np.uint16)
                if no_pl.any():
                    hard = wei[no_pl]
                    codes[no_pl] = np.where(hard == 1.0, 0, top).astype(np.uint16)
                if np.array_equal(table[codes.astype(np.int64)], wei):
                    self._pl_codes = (codes, table)
        else:
            wei = self.get_wei_from_GT(gt)
        return (v['chr'][req], v['pos'][req], gt, wei, v['dp'][req])

    def filter_chr_names(self):
        """parsers.py:159-163: strip every 'chr' (any case); ids in first-appearance order."""
        from . import labels
        self.g_chrs, _, mapped = labels.map_labels(self.chrs, lambda c: re.sub("chr", "", c, flags=re.IGNORECASE))
        self.g_chrs_ids = labels.factorize(mapped)[1] if len(mapped) else self.g_chrs

    def save_to_bed(self, outFile):
        with open(outFile, "w") as fh:
            for c, p, g in zip(self.chrs, self.pos, self.gt):
                fh.write("%s\t%s\t%s\n" % (c, p, g))


def potatoParser(inFile, logDebug, outFile="parser"):
    inputs = ParseInputs(inFile, logDebug, outFile)
    return (inputs.chrs, inputs.pos, inputs.gt, inputs.wei, inputs.dp)
