"""
`snpmatch simulate` on the resident panel — mirrors `snpmatch/core/simulate.py`:
`simulateSNPs` (:10-31), `simulateSNPs_F1` (:33-60), `potatoSimulate` (:62-68).

The reference reads whole accession columns from its second, column-chunked HDF5 file
(`g.g_acc.snps[:, ix]`, simulate.py:15,36-37); here the columns come out of the one HBM-resident
2-bit panel with a column kernel (`snpm_db_read_columns`).  The random draws use `np.random` in the
reference's call order, so that the same `np.random.seed` gives the same markers, errors and hets.

Deliberate departures (the reference cannot run as written under Python 3): accession ids are
compared as text (`g.accessions`; the reference compares `str` with the `bytes` array
`g.g.accessions`, simulate.py:12-13,34-35, which never matches), and the genotype column is written as
text (`0/0`), not as the repr of `bytes` (`b'0/0'`, simulate.py:28,57).
"""
import logging

import numpy as np
import pandas as pd

from . import parsers
from . import snp_genotype

log = logging.getLogger(__name__)


def _row_chromosomes(panel, rows):
    """Chromosome label of database rows without building the N-string list of pygwas/genotype.py:156-161."""
    starts = np.asarray(panel.chr_regions)[:, 0]
    return np.asarray(panel.chrs).astype("U")[np.searchsorted(starts, rows, side="right") - 1]


def _frame(chrs, pos, snp):
    return pd.DataFrame({"chr": np.asarray(chrs), "pos": np.asarray(pos), "snp": np.asarray(snp)})


def _finish(input_df, outFile):
    input_df["snp"] = parsers.snp_binary_to_gt(np.array(input_df["snp"], dtype="int8")).astype("U")
    if outFile is not None:
        input_df.to_csv(outFile, sep="\t", index=None, header=False)
    return input_df


def simulateSNPs(g, AccID, numSNPs, outFile=None, err_rate=0.001):
    assert type(AccID) is str, "provide Accession ID as a string"
    assert AccID in g.accessions, "accession is not present in the matrix!"
    AccToCheck = np.where(g.accessions == AccID)[0][0]
    log.info("reading the column of accession %s from the resident panel", AccID)
    acc_snp = g.g_acc.snps[:, AccToCheck]
    informative_snps = np.where(acc_snp >= 0)[0]            # removing NAs for the accession
    log.info("drawing %s of %s called positions", numSNPs, informative_snps.shape[0])
    sampleSNPs = np.sort(np.random.choice(np.arange(informative_snps.shape[0]), numSNPs, replace=False))
    rows = informative_snps[sampleSNPs]
    snp = acc_snp[rows].astype(np.int8)
    num_to_change = int(err_rate * numSNPs)
    log.info("replacing %s calls by random ones (error rate %s)", num_to_change, err_rate)
    new_calls = np.random.choice(3, num_to_change)          # drawn first: simulate.py:26 evaluates its right-hand side first
    change = np.sort(np.random.choice(np.arange(numSNPs), num_to_change, replace=False))
    snp[change] = new_calls
    return _finish(_frame(_row_chromosomes(g.g, rows), g.g.positions[rows], snp), outFile)


def simulateSNPs_F1(g, parents, numSNPs, outFile, err_rate, rm_hets=1):
    indP1 = np.where(g.accessions == parents.split("x")[0])[0][0]
    indP2 = np.where(g.accessions == parents.split("x")[1])[0][0]
    log.info("reading the parents' columns from the resident panel")
    cols = g.g_acc.snps[:, [indP1, indP2]]
    snpsP1, snpsP2 = cols[:, 0], cols[:, 1]
    common_ix = np.where((snpsP1 >= 0) & (snpsP2 >= 0) & (snpsP1 < 2) & (snpsP2 < 2))[0]
    common_snps = np.where(snpsP1[common_ix] != snpsP2[common_ix], 2, snpsP1[common_ix]).astype("int8")
    log.info("drawing %s of %s positions called homozygous in both parents", numSNPs, common_ix.shape[0])
    sampleSNPs = np.sort(np.random.choice(np.arange(common_ix.shape[0]), numSNPs, replace=False))
    rows = common_ix[sampleSNPs]
    snp = common_snps[sampleSNPs].astype(int)
    num_to_change = int(err_rate * numSNPs)
    log.info("replacing %s homozygous calls by random ones (error rate %s)", num_to_change, err_rate)
    new_calls = np.random.choice(2, num_to_change)          # simulate.py:52: right-hand side first
    change = np.sort(np.random.choice(np.where(snp != 2)[0], num_to_change, replace=False))
    snp[change] = new_calls
    # also change hets randomly to homozygous
    het_ix = np.where(snp == 2)[0]
    snp[het_ix] = np.random.choice(3, het_ix.shape[0], p=[(1 - rm_hets) / 2, (1 - rm_hets) / 2, rm_hets])
    return _finish(_frame(_row_chromosomes(g.g_acc, rows), np.asarray(g.g_acc.positions)[rows], snp), outFile)


def potatoSimulate(args):
    g = snp_genotype.Genotype(args['hdf5File'], args['hdf5accFile'])
    if args['simF1']:
        simulateSNPs_F1(g, args['AccID'], args['numSNPs'], args['outFile'], args['err_rate'], args['rm_het'])
    else:
        simulateSNPs(g, args['AccID'], args['numSNPs'], args['outFile'], args['err_rate'])
    log.info("finished!")
