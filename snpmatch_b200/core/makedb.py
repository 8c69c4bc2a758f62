"""
`snpmatch makedb` — build the database this package matches against.

Mirrors the steps of `snpmatch/core/makedb.py` without its external tools: the reference shells out to bcftools + sed to turn
a multi-sample VCF into `Chromosome,Position,<accession>...` rows of 0 / 1 / 2 / -1 (`getCSV`, makedb.py:34-62), parses that
CSV into per-chromosome lists (`pygwas/genotype.py:29-105`) and writes two HDF5 files, one chunked by row and one by column
(`makeHDF5s`, makedb.py:83-90).  Here the VCF is read directly (`vcf_to_csv`), the CSV is parsed with one vectorised read
(`load_csv`), and the matrix goes to the GPU once, where it is packed to 2 bits per call (`k_pack_int8`); the packed words are
written as ONE native file (`<db_id>.npz`, `Genotype.save_packed`) that serves row and column access and is what
`-d` of every other sub-command loads.  The genome JSON of the contig lengths (`<db_id>.json`) is written as in the reference.

Departure: genotype calls other than 0/0, 1/1, 0/1, 1/0, ./. (e.g. a third allele) pass through the reference's sed chain
unchanged and then break its CSV parser; here they are stored as missing (-1) and counted in the log.
"""
import gzip
import json
import logging
import os.path
import re
import sys

import numpy as np
import pandas as pd

from . import snp_genotype

log = logging.getLogger(__name__)

_GT_CODE = {"0/0": "0", "0|0": "0", "1/1": "1", "1|1": "1", "0/1": "2", "0|1": "2", "1/0": "2", "1|0": "2", "./.": "-1", ".|.": "-1", ".": "-1"}


def die(msg):
    sys.stderr.write('Error: ' + msg + '\n')
    sys.exit(1)


def get_contigs(vcf_header):
    """Chromosome names and lengths from the ##contig lines of a VCF header (makedb.py:25-32)."""
    chr_names, chr_len = [], []
    for eh in vcf_header:
        if eh[0:8] == '##contig':
            body = eh.replace(">", "").replace("<", "")
            chr_names.append(body.split("ID=")[1].split(",")[0])
            chr_len.append(int(body.split("length=")[1].split(",")[0]))
    return {"ref_chrs": chr_names, "ref_chrlen": chr_len}


def vcf_to_csv(inVCF, outFile):
    """What getCSV (makedb.py:34-62) produces with bcftools + sed, read straight from the VCF: `<outFile>.csv` with the header
    `Chromosome,Position,<samples>` and one row of 0 / 1 / 2 / -1 per record, and `<outFile>.json` with the contigs."""
    opener = gzip.open if str(inVCF).endswith(".gz") else open
    header, other = [], 0
    with opener(inVCF, "rt") as fh, open(outFile + ".csv", "w") as out:
        for line in fh:
            if line.startswith("##"):
                header.append(line.rstrip("\n"))
                continue
            f = line.rstrip("\n").split("\t")
            if line.startswith("#"):
                out.write("Chromosome,Position" + "".join("," + s for s in f[9:]) + "\n")
                with open(outFile + ".json", "w") as js:
                    js.write(json.dumps(get_contigs(header), sort_keys=True, indent=4))
                continue
            if len(f) < 10:
                continue
            keys = f[8].split(":")
            at = keys.index("GT") if "GT" in keys else -1
            codes = []
            for cell in f[9:]:
                v = cell.split(":")
                code = _GT_CODE.get(v[at] if 0 <= at < len(v) else ".")
                if code is None:
                    code = "-1"
                    other += 1
                codes.append(code)
            out.write(f[0] + "," + f[1] + "".join("," + c for c in codes) + "\n")
    if other:
        log.warning("%s genotype calls are neither 0/0, 1/1, 0/1 nor ./. and were stored as missing", other)
    log.info("Number of contigs found: %s", len(get_contigs(header)["ref_chrs"]))


def _sniff(csvFile):
    with open(csvFile) as fh:
        head = fh.readline().rstrip("\n")
    sep = "\t" if ("\t" in head and "," not in head) else ","
    cols = [c.strip() for c in head.split(sep)]
    if len(cols) < 2 or cols[0] != 'Chromosome' or cols[1] not in ('Positions', 'Position'):
        raise Exception('First two columns must be in form  Chromosome, Positions')
    return sep, cols


def _index_arrays(csvFile, sep):
    """(chrs, chr_regions, positions) of pygwas' `load_csv_genotype_data` (pygwas/genotype.py:29-105): a chromosome entry per run
    of equal labels in file order, `chr_regions` the row ranges of the runs.  Reads the first two columns only."""
    idx = pd.read_csv(csvFile, sep=sep, usecols=[0, 1], dtype=str)
    labels = idx.iloc[:, 0].to_numpy().astype("U")
    positions = idx.iloc[:, 1].to_numpy().astype(np.int64).astype(np.int32)
    n = len(labels)
    starts = np.concatenate([[0], np.flatnonzero(labels[1:] != labels[:-1]) + 1]) if n else np.zeros(0, dtype=np.int64)
    ends = np.concatenate([starts[1:], [n]]) if n else starts
    regions = np.stack([starts, ends], axis=1).astype(np.int64) if n else np.zeros((0, 2), dtype=np.int64)
    return (labels[starts] if n else labels), regions, positions


def load_csv(csvFile):
    """All arrays of the CSV in memory (small files, tests): snps int8 [N,A], positions, chrs, chr_regions, accessions (bytes)."""
    sep, cols = _sniff(csvFile)
    chrs, regions, positions = _index_arrays(csvFile, sep)
    geno = pd.read_csv(csvFile, sep=sep, usecols=range(2, len(cols)), dtype=np.int8) if len(cols) > 2 else pd.DataFrame(index=range(len(positions)))
    return {"snps": np.ascontiguousarray(geno.to_numpy(dtype=np.int8)).reshape(len(positions), len(cols) - 2), "positions": positions,
            "chrs": chrs, "chr_regions": regions, "accessions": np.array(cols[2:], dtype="S")}


def makeDB(csvFile, outFile, device=0, chunk_rows=200000):
    """CSV -> resident 2-bit panel -> `<outFile>.npz` (replaces makeHDF5s, makedb.py:83-90: one native file instead of the
    row-chunked and the column-chunked HDF5).  The genotype columns are streamed to the GPU `chunk_rows` rows at a time, so
    the int8 matrix (12 GB for the 1001-genomes panel) never sits in host memory.  Returns the open Genotype."""
    from .. import lib
    sep, cols = _sniff(csvFile)
    chrs, regions, positions = _index_arrays(csvFile, sep)
    n_acc = len(cols) - 2
    assert n_acc > 0, "the CSV holds no accession columns"
    log.info("packing %s SNPs x %s accessions on the GPU", len(positions), n_acc)
    db = lib.Database(positions, regions, n_acc, device=device)
    row0 = 0
    for block in pd.read_csv(csvFile, sep=sep, usecols=range(2, len(cols)), dtype=np.int8, chunksize=int(chunk_rows)):
        codes = np.ascontiguousarray(block.to_numpy(dtype=np.int8))
        db.load_int8(codes, row0=row0)
        row0 += len(codes)
    assert row0 == len(positions), "the CSV changed while it was read"
    g = object.__new__(snp_genotype.Genotype)
    g._finish(db, positions, chrs, regions, np.array(cols[2:], dtype="S"))
    g.save_packed(outFile + ".npz")
    log.info("wrote %s.npz", outFile)
    return g


def makedb_from_vcf(args):
    _, inType = os.path.splitext(args['inFile'])
    if inType == '.vcf' or len(re.compile(".vcf.gz$").findall(os.path.basename(args['inFile']))) > 0:
        log.info("VCF -> %s.csv", args['db_id'])
        vcf_to_csv(args['inFile'], args['db_id'])
        makeDB(args['db_id'] + '.csv', args['db_id']).close()
    elif inType == '.csv':
        makeDB(args['inFile'], args['db_id']).close()
    else:
        die("please provide either a VCF file or a CSV!")
