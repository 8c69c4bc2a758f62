"""
TEST INFRASTRUCTURE — not product code.

Loads the UNMODIFIED reference (read-only, /root/reference) in this container so
that golden vectors can be generated from the reference's own functions
(SURVEY.md Appendix B).  Only `oracle/gen_golden.py` and tests that are skipped
when /root/reference is absent may import this module.  Nothing here travels
to the GPU box: the vectors it produces are committed under tests/golden/.

Shims (none of them touches hot-path arithmetic):
  * stub modules `allel`, `h5py`, `hmmlearn(.hmm)` — imported at module top by the
    reference (parsers.py:3, snp_genotype.py:12, pygwas/genotype.py:4, infer.py:8)
    but never called on the scoring path;
  * `pandas.DataFrame.append` -> `pd.concat` (csmatch.py:91 uses the removed API);
  * a NumPy-backed stand-in for `HDF5Genotype` (pygwas/genotype.py:534-673).
"""
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("SNPMATCH_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "snpmatch", "core"))


def load_reference():
    """Import the reference package; returns a namespace of its hot-path modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("allel", "h5py", "hmmlearn", "hmmlearn.hmm"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["hmmlearn"].hmm = sys.modules["hmmlearn.hmm"]
    import pandas as pd
    if not hasattr(pd.DataFrame, "append"):
        def _append(self, other, ignore_index=False):
            if len(self) == 0:
                return other.reset_index(drop=True) if ignore_index else other
            return pd.concat([self, other], ignore_index=ignore_index)
        pd.DataFrame.append = _append
    if not hasattr(np, "in1d"):
        np.in1d = np.isin
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    warnings.filterwarnings("ignore", category=FutureWarning)
    import logging
    logging.getLogger().setLevel(logging.ERROR)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        from snpmatch.core import snpmatch as r_snpmatch
        from snpmatch.core import csmatch as r_csmatch
        from snpmatch.core import snp_genotype as r_snp_genotype
        from snpmatch.core import parsers as r_parsers
        from snpmatch.core import genomes as r_genomes
    return types.SimpleNamespace(snpmatch=r_snpmatch, csmatch=r_csmatch,
                                 snp_genotype=r_snp_genotype, parsers=r_parsers,
                                 genomes=r_genomes)


class FakeHDF5Genotype(object):
    """NumPy stand-in for pygwas.genotype.HDF5Genotype (pygwas/genotype.py:534-673)."""

    def __init__(self, snps, positions, chrs, chr_regions, accessions):
        self.snps = np.ascontiguousarray(snps, dtype=np.int8)
        self.positions = np.asarray(positions, dtype=np.int32)
        self.chrs = np.asarray(chrs, dtype="str")
        self.chr_regions = np.asarray(chr_regions, dtype=np.int64)
        self.accessions = np.asarray(accessions, dtype="S")

    @property
    def chromosomes(self):
        # pygwas/genotype.py:156-161 — one chromosome label per SNP row
        out = []
        for c, (s, e) in zip(self.chrs, self.chr_regions):
            out.extend([str(c)] * int(e - s))
        return out


def make_reference_genotype(ref, snps, positions, chrs, chr_regions, accessions):
    """Wrap arrays as the reference's snp_genotype.Genotype (snp_genotype.py:26-41)."""
    db = FakeHDF5Genotype(snps, positions, chrs, chr_regions, accessions)
    G = object.__new__(ref.snp_genotype.Genotype)
    G.g = db
    G.g_acc = db
    G.accessions = db.accessions.astype("U")
    G.chrs = db.chrs.astype("U")
    return G


def make_reference_inputs(ref, chrs, pos, gt, wei, dp):
    """ParseInputs via load_snp_info (parsers.py:89-94), bypassing file I/O."""
    inp = ref.parsers.ParseInputs("")
    inp.load_snp_info(chrs, pos, gt, wei, dp)
    return inp
