"""
Batched many-sample mode (north-star item (d)).

The reference has no batched entry point: many samples are genotyped as one `snpmatch inbred` process per sample
(README.md:9).  When the samples are CALLED genotypes (BED files, VCFs without PL: one-hot weights,
parsers.py:132-139) they can instead be scored together: their matched database rows are merged into one shared marker
panel, every sample becomes a row of 2-bit calls on that panel, and scores and informative-site counts of all samples
against all accessions come out of ONE one-hot int8 GEMM on the tensor cores (csrc/gemm_onehot.cuh).  Results equal
those of `Genotyper.genotyper` run sample by sample (integers, exactly).
"""
import numpy as np

from .. import lib
from . import parsers
from . import snpmatch


def shared_panel(g, samples):
    """Join every sample with the database (GPU) and merge the matched rows.
    Returns (panel_rows int64[K] global rows ascending, codes uint8 [S,K] with 3 = sample lacks the marker,
    matched list of (db_idx, sample_idx) per sample)."""
    matched = []
    for inp in samples:
        matched.append(g.get_positions_idxs(inp.chrs, inp.pos))
    rows = np.unique(np.concatenate([m[0] for m in matched])) if matched else np.zeros(0, dtype=np.int64)
    codes = np.full((len(samples), len(rows)), 3, dtype=np.uint8)
    for i, (inp, (db_idx, s_idx)) in enumerate(zip(samples, matched)):
        c = parsers.parseGT(np.asarray(inp.gt)[s_idx])            # 0 ref, 1 alt, 2 het, -1 no call
        col = np.searchsorted(rows, db_idx)
        codes[i, col] = np.where(c < 0, 3, c).astype(np.uint8)
    return rows.astype(np.int64), codes, matched


def genotype_batch(g, samples, skip_db_hets=False):
    """Score S called-genotype samples in one tensor-core pass; returns (list of GenotyperOutput, info dict).
    Every sample's weights must be one-hot (use Genotyper for likelihood-weighted samples)."""
    for inp in samples:
        assert lib.weights_are_one_hot(inp.wei), "genotype_batch scores called genotypes; use Genotyper for PL-weighted samples"
    rows, codes, matched = shared_panel(g, samples)
    local_rows = rows                                             # score_shared_panel takes global rows
    r = g.db.score_shared_panel(local_rows, codes, skip_db_hets=skip_db_hets)
    out = []
    for i, (inp, (db_idx, _)) in enumerate(zip(samples, matched)):
        m = len(db_idx)
        res = snpmatch.GenotyperOutput(g.g.accessions, r["matches"][i], r["ninfo"][i], snpmatch.get_fraction(m, len(inp.pos)), m, inp.dp)
        res._attach_fused(r["prob"][i], r["L"][i], r["LR"][i])
        out.append(res)
    return out, {"panel_markers": len(rows), "gemm_ms": r["gemm_ms"]}


def genotype_many(g, samples, skip_db_hets=False):
    """Genotyper.genotyper (snpmatch.py:207-233) for MANY samples with any weights (PL likelihoods or called genotypes) in
    one device pass: the markers of every sample are ordered by weight triple (host, `lib.group_markers`; what a parser
    would cache next to <input>.snpmatch.npz) and scored by the counting kernel (csrc/grouped.cuh).  Returns the list of
    GenotyperOutput, one per sample, with matches / ninfo / overlap identical to Genotyper run sample by sample and
    likelihoods within 1e-9; samples whose int(score) would depend on the reference's summation order are re-scored by
    the order-exact kernel inside `lib.score_grouped`."""
    prepared = [g.prepare_markers(inp.chrs, inp.pos) for inp in samples]
    offs = np.concatenate([[0], np.cumsum([len(p[2]) for p in prepared])]).astype(np.int64)
    cid = np.concatenate([p[1] for p in prepared]) if samples else np.zeros(0, np.int32)
    pos = np.concatenate([p[2] for p in prepared]) if samples else np.zeros(0, np.int32)
    wei = np.concatenate([np.asarray(inp.wei, dtype=np.float64)[p[0]] for inp, p in zip(samples, prepared)]) if samples else np.zeros((0, 3))
    r = lib.score_grouped(g.db, offs, cid, pos, wei, skip_db_hets=skip_db_hets)
    out = []
    for i, inp in enumerate(samples):
        m = int(r["m"][i])
        res = snpmatch.GenotyperOutput(g.g.accessions, r["score"][i], r["ninfo"][i], snpmatch.get_fraction(m, len(inp.pos)), m, inp.dp)
        res._attach_fused(r["prob"][i], r["L"][i], r["LR"][i])
        out.append(res)
    return out
