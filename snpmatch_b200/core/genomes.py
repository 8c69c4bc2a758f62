"""
Genome / window geometry for `snpmatch cross` (host side).

Mirrors the interface of the reference's `snpmatch/core/genomes.py`: `Genome` (genomes.py:16-70) and
the window iterators `get_bins_genome` / `get_bins_arrays` / `get_bins_echr` (genomes.py:73-127).
Window k of a chromosome covers [1 + k*b, (k+1)*b] for k = 0.. while 1 + k*b < chrlen
(genomes.py:113-116); exactly one tuple per window is produced, empty windows included.

On the hot path the windows are not iterated at all: `window_layout` turns the genome JSON into
per-database-chromosome (count, first index) pairs and the device assigns every matched marker
its window number (csrc/windows.cuh).  The iterators remain for API compatibility; they are
vectorised with searchsorted instead of the reference's per-position Python loop.
"""
import glob
import json
import logging
import os.path

import numpy as np

log = logging.getLogger(__name__)

_RES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "resources", "genomes")


def genome_style_ids(names):
    """lower-case, then drop 'chr' (genomes.py:28,75,95)."""
    from . import labels
    return labels.map_labels(names, lambda c: c.lower().replace("chr", ""))[0]


def num_windows(chrlen, bin_len):
    return len(range(1, int(chrlen), int(bin_len)))


class Genome(object):
    """Chromosome names and lengths of a reference assembly, from a JSON file or a bundled id."""

    def __init__(self, ref_json):
        if ref_json in self.get_genome_ids():
            ref_json = os.path.join(_RES, ref_json + ".json")
        assert os.path.exists(ref_json), "Reference json file missing: %s" % ref_json
        with open(ref_json) as fh:
            self.json = json.load(fh)
        self.chrs = np.array(self.json["ref_chrs"], dtype="str")
        self.chrlen = np.array(self.json["ref_chrlen"], dtype=int)
        self.chrs_ids = genome_style_ids(self.chrs)

    @staticmethod
    def get_genome_ids():
        return [os.path.basename(f)[:-len(".json")] for f in glob.glob(os.path.join(_RES, "*.json"))]

    def get_chr_ind(self, echr):
        """Index of a chromosome name (or of every name of an array) in the genome (genomes.py:38-51)."""
        real = np.array([c.replace("Chr", "").replace("chr", "") for c in self.chrs])
        if isinstance(echr, (str, bytes, np.str_, np.bytes_)):
            name = echr.decode() if isinstance(echr, bytes) else str(echr)
            hit = np.flatnonzero(real == name.replace("Chr", "").replace("chr", ""))
            return int(hit[0]) if len(hit) == 1 else None
        echr = np.asarray(echr)
        out = np.zeros(len(echr), dtype="int8")
        for name in np.unique(echr):
            hit = np.flatnonzero(real == str(name).replace("Chr", "").replace("chr", ""))
            out[echr == name] = hit[0]
        return out

    def estimated_cM_distance(self, snp_position):
        """cM estimate from the per-chromosome mean recombination rate (genomes.py:53-70)."""
        rates = self.json.get("recomb_rates")
        if rates is None:
            log.warning("no 'recomb_rates' in the genome json; using 3 cM/Mb")
            rates = np.repeat(3, len(self.chrs_ids))
        assert isinstance(snp_position, str), "expected a string!"
        f = snp_position.split(",")
        assert len(f) >= 2, "input should be 'chr1,1000' or 'chr1,1000,2000'"
        where = int(f[1]) if len(f) == 2 else (int(f[1]) + int(f[2])) / 2
        return rates[self.get_chr_ind(f[0])] * where / 1000000

    # ---- hot-path geometry ---------------------------------------------------------------------
    def window_layout(self, db_chrs, bin_len):
        """For a database chromosome list: (win_count[C_db], win_off[C_db], n_windows, winds_chrs).

        win_count[c] = number of windows of database chromosome c (0 when it is not in the genome),
        win_off[c] = 0-based index of its first window in genome-JSON order; winds_chrs = genome id of
        every window (CrossIdentifier.window_genotyper's `winds_chrs`, csmatch.py:94)."""
        bin_len = int(bin_len)
        db_ids = genome_style_ids(db_chrs)
        assert len(db_ids) <= len(self.chrs_ids), "Please change default --genome option"
        assert len(np.intersect1d(db_ids, self.chrs_ids)) > 0, "Please change default --genome option"
        if len(np.intersect1d(db_ids, self.chrs_ids)) < len(self.chrs_ids):
            log.warning("Some reference contigs are missing in genotype hdf5 file")
        counts = np.array([num_windows(l, bin_len) for l in self.chrlen], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(counts)])
        win_count = np.zeros(len(db_ids), dtype=np.int32)
        win_off = np.zeros(len(db_ids), dtype=np.int32)
        for gi, gid in enumerate(self.chrs_ids):
            hit = np.flatnonzero(db_ids == gid)
            if len(hit):                      # the reference takes the first database chromosome with this id
                win_count[hit[0]] = counts[gi]
                win_off[hit[0]] = offs[gi]
        winds_chrs = np.repeat(self.chrs_ids, counts)
        return win_count, win_off, int(offs[-1]), winds_chrs

    # ---- iterator API of the reference -----------------------------------------------------------
    def get_bins_genome(self, g, binLen):
        binLen = int(binLen)
        g_ids = genome_style_ids(g.chrs)
        assert len(g_ids) <= len(self.chrs_ids), "Please change default --genome option"
        assert len(np.intersect1d(g_ids, self.chrs_ids)) > 0, "Please change default --genome option"
        for chr_ix, cid in enumerate(self.chrs_ids):
            hit = np.flatnonzero(g_ids == cid)
            if len(hit):
                start, end = int(g.chr_regions[hit[0]][0]), int(g.chr_regions[hit[0]][1])
                chr_pos = np.asarray(g.positions[start:end])
            else:                              # a contig missing from the database is empty
                start, chr_pos = 0, np.zeros(0, dtype=int)
            for e_bin in get_bins_echr(self.chrlen[chr_ix], chr_pos, binLen, start):
                yield (chr_ix, e_bin[0], e_bin[1])

    def get_bins_arrays(self, g_chrs, g_snppos, binLen):
        binLen = int(binLen)
        ids = genome_style_ids(g_chrs)
        uniq = np.unique(ids)
        assert len(uniq) <= len(self.chrs_ids), "Please change default --genome option"
        assert len(np.intersect1d(uniq, self.chrs_ids)) > 0, "Please change default --genome option"
        g_snppos = np.asarray(g_snppos)
        for chr_ix, cid in enumerate(self.chrs_ids):
            ix = np.flatnonzero(ids == cid)
            rel = int(ix[0]) if len(ix) else 0
            for e_bin in get_bins_echr(self.chrlen[chr_ix], g_snppos[ix], binLen, rel):
                yield (chr_ix, e_bin[0], e_bin[1])


def get_bins_echr(real_chrlen, chr_pos, binLen, rel_ix):
    """Yield ([start, end], [indices]) per window; positions ascending (genomes.py:111-127)."""
    chr_pos = np.asarray(chr_pos)
    binLen = int(binLen)
    starts = np.arange(1, int(real_chrlen), binLen, dtype=np.int64)
    lo = np.searchsorted(chr_pos, starts, side="left")
    hi = np.searchsorted(chr_pos, starts + binLen - 1, side="right")
    for t, a, b in zip(starts, lo, hi):
        yield ([int(t), int(t) + binLen - 1], list(range(int(a) + rel_ix, int(b) + rel_ix)))
