"""
TEST INFRASTRUCTURE — CPU oracle for the genotype-matching hot path.

A NumPy restatement of the reference algorithm (Gregor-Mendel-Institute/SNPmatch
5.0.1) for the path SURVEY.md section 8 names.  It is the CHECKER for the CUDA
path: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The product (`snpmatch_b200`) never does.

Pinning: every function below is checked bit-for-bit (integers, and the fp64
scores) against the UNMODIFIED reference run in the build container — see
`oracle/gen_golden.py` (generator), `tests/golden/` (committed vectors) and
`tests/test_oracle_golden.py`; and against the reference's own known answers
(`likeliTest(10,3) == 122.8361221819443`, tests/test_inbred.py:22).

All `file:line` citations are relative to the reference tree.
"""
import itertools
import re

import numpy as np

LR_THRES = 3.841          # snpmatch.py:17
SNP_THRES = 4000          # snpmatch.py:18
PROB_THRES = 0.98         # snpmatch.py:19
CHUNK_SIZE = 1000         # snpmatch.py:173
P_MATCH = 0.99999999      # snpmatch.py:44

_CHR_RE = re.compile("chr", re.IGNORECASE)


# --------------------------------------------------------------------------
# A1 — (chrom, pos) join
# --------------------------------------------------------------------------
def normalize_chr_names(chrs):
    """parsers.py:161 — delete every 'chr' (any case) from each label."""
    return np.array([_CHR_RE.sub("", str(c)) for c in np.asarray(chrs).ravel()], dtype="str")


def first_appearance_ids(labels):
    """parsers.py:162-163 — unique labels in first-appearance order."""
    _, first = np.unique(labels, return_index=True)
    return labels[np.sort(first)]


def get_common_positions(chr1, pos1, chr2, pos2):
    """snp_genotype.py:46-68.  Side 1 is the database, side 2 the sample.

    Per chromosome, in side-1 first-appearance order, membership of each side's
    positions in the other (`np.in1d(..., assume_unique=True)`), indices appended.
    """
    pos1 = np.asarray(pos1)
    pos2 = np.asarray(pos2)
    assert len(chr1) == len(pos1) and len(chr2) == len(pos2)
    g1 = normalize_chr_names(chr1)
    g2 = normalize_chr_names(chr2)
    ids1 = first_appearance_ids(g1) if len(g1) else g1
    ids2 = set(first_appearance_ids(g2).tolist()) if len(g2) else set()
    out1, out2 = [], []
    for cid in ids1:
        if cid not in ids2:
            continue
        ix1 = np.flatnonzero(g1 == cid)
        ix2 = np.flatnonzero(g2 == cid)
        p1 = pos1[ix1].astype(np.int64)
        p2 = pos2[ix2].astype(np.int64)
        out1.append(ix1[np.isin(p1, p2, assume_unique=True)])
        out2.append(ix2[np.isin(p2, p1, assume_unique=True)])
    if not out1:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    return np.concatenate(out1).astype(np.int64), np.concatenate(out2).astype(np.int64)


def db_chromosome_labels(chrs, chr_regions):
    """pygwas/genotype.py:156-161 — one label per database row."""
    reps = [int(e) - int(s) for s, e in chr_regions]
    return np.repeat(np.asarray(chrs, dtype="str"), reps)


# --------------------------------------------------------------------------
# A2 — matchGTsAccs
# --------------------------------------------------------------------------
def match_gts_accs(sample_wei, db_snps, skip_hets_db=False):
    """snpmatch.py:74-89.

    db codes: 0 hom-ref, 1 hom-alt, 2 het, <0 missing.  Weight column mapping
    (snpmatch.py:81-87): ref -> wei[:,0], het -> wei[:,1], alt -> wei[:,2].
    Summation order: per class a plain left-to-right sum over the
    rows, classes combined as ((0 + ref) + het) + alt.  Reducing a C-ordered
    (k, A) product over axis 0 adds row vectors one after the other, which is that
    order exactly (no pairwise tree on a strided axis).
    """
    sample_wei = np.asarray(sample_wei, dtype=np.float64)
    db = np.array(db_snps, dtype=np.int8, copy=True)
    assert sample_wei.shape[0] == db.shape[0]
    assert sample_wei.ndim == 2 and sample_wei.shape[1] == 3
    if skip_hets_db:
        db[db == 2] = -1                      # snpmatch.py:78-79
    n_acc = db.shape[1]
    score = np.zeros(n_acc, dtype=np.float64)
    for code, col in ((0, 0), (2, 1), (1, 2)):
        prod = (db == code).astype(np.float64) * sample_wei[:, col][:, None]
        score = score + np.add.reduce(prod, axis=0)
    ninfo = (db >= 0).sum(axis=0).astype(np.int64)
    return score, ninfo


def match_gts_accs_sequential(sample_wei, db_snps, skip_hets_db=False):
    """The same arithmetic as an explicit row-by-row scalar recurrence — the order the
    CUDA kernel implements.  Slow; used to pin `match_gts_accs` on small inputs."""
    sample_wei = np.asarray(sample_wei, dtype=np.float64)
    db = np.array(db_snps, dtype=np.int8, copy=True)
    if skip_hets_db:
        db[db == 2] = -1
    n_acc = db.shape[1]
    s_ref = np.zeros(n_acc)
    s_het = np.zeros(n_acc)
    s_alt = np.zeros(n_acc)
    ninfo = np.zeros(n_acc, dtype=np.int64)
    for k in range(db.shape[0]):
        row = db[k]
        s_ref[row == 0] += sample_wei[k, 0]
        s_het[row == 2] += sample_wei[k, 1]
        s_alt[row == 1] += sample_wei[k, 2]
        ninfo += (row >= 0)
    return ((np.zeros(n_acc) + s_ref) + s_het) + s_alt, ninfo


# --------------------------------------------------------------------------
# A4 — likelihood epilogue
# --------------------------------------------------------------------------
def get_fraction(x, y, y_min=0):
    """snpmatch.py:25-28."""
    if y <= y_min:
        return np.nan
    return float(x) / y


def likeli_test(n, y):
    """snpmatch.py:40-55 — n informative sites, y matched (may be float)."""
    assert y <= n, "provided y is greater than n"
    if n == 0:
        return np.nan
    if y == n:
        return 1.0
    if y > 0:
        p_s = float(y) / n
        return y * np.log(p_s / P_MATCH) + (n - y) * np.log((1 - p_s) / (1 - P_MATCH))
    return np.nan


def calculate_likelihoods(scores, ninfo, amin="calc"):
    """snpmatch.py:106-117 — vectorised restatement; returns (L, LR)."""
    y = np.asarray(scores, dtype=np.float64)
    n = np.asarray(ninfo, dtype=np.float64)
    assert np.all(y <= n), "provided y is greater than n"
    with np.errstate(divide="ignore", invalid="ignore"):
        p_s = y / n
        lik = y * np.log(p_s / P_MATCH) + (n - y) * np.log((1 - p_s) / (1 - P_MATCH))
    lik = np.where(y == n, 1.0, lik)
    lik = np.where((n == 0) | (y <= 0), np.nan, lik)
    if amin == "calc":
        top = np.nan if np.all(np.isnan(lik)) else np.nanmin(lik)
    else:
        top = float(amin)
    if not (top > 0):
        lr = np.full(lik.shape, np.nan)       # get_fraction(L, top): top <= 0 -> nan; nan top -> x/nan
    else:
        lr = lik / top
    return lik, lr


def probabilities(scores, ninfo):
    """snpmatch.py:102-104."""
    y = np.asarray(scores, dtype=np.float64)
    n = np.asarray(ninfo, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = y / n
    return np.where(n <= 0, np.nan, p)


def test_identity(x, n, error_rate=0.0005, pthres=0.05):
    """snpmatch.py:57-72 — binom.sf((n - x) - 1, n, error_rate) >= pthres."""
    from scipy import stats
    x = np.asarray(x, dtype=np.float64)
    n = np.asarray(n)
    st = stats.binom.sf(n - x - 1, n, error_rate)
    return np.array(st >= pthres).astype(int)


test_identity.__test__ = False  # not a pytest test


def identity_kmax_table(n_max, error_rate=0.02, pthres=0.05):
    """kmax[n] = largest integer k with binom.sf(k - 1, n, e) >= pthres, so that
    identical <=> floor(n - x - 1) + 1 <= kmax[n] (SciPy floors sf's first argument and sf is
    monotone in it); built with the same SciPy call as `test_identity`.  kmax >= 0 (sf(-1)=1)."""
    from scipy import stats
    kmax = np.zeros(n_max + 1, dtype=np.int32)
    for n in range(n_max + 1):
        ks = np.arange(0, n + 2)
        ok = stats.binom.sf(ks - 1, n, error_rate) >= pthres
        kmax[n] = int(np.flatnonzero(ok).max())
    return kmax


# --------------------------------------------------------------------------
# A3 — Genotyper.genotyper
# --------------------------------------------------------------------------
class InbredResult(object):
    """GenotyperOutput fields (snpmatch.py:94-100) plus the untruncated float scores."""

    def __init__(self, score_f64, ninfo, overlap, num_snps, common):
        self.score_f64 = score_f64
        self.scores = np.array(score_f64, dtype="int")        # truncation, snpmatch.py:96
        self.ninfo = np.array(ninfo, dtype="int")
        self.overlap = overlap
        self.num_snps = num_snps
        self.common = common
        self.probabilities = probabilities(self.scores, self.ninfo)
        self.likelis, self.lrts = calculate_likelihoods(self.scores, self.ninfo)


def genotyper(db_snps, db_chrs, db_chr_regions, db_positions, s_chrs, s_pos, s_wei,
              skip_db_hets=False, chunk_size=CHUNK_SIZE, filter_pos_ix=None):
    """snpmatch.py:207-233 — join, then 1000-row chunks of matchGTsAccs accumulated in order."""
    labels = db_chromosome_labels(db_chrs, db_chr_regions)
    common = get_common_positions(labels, db_positions, s_chrs, s_pos)
    if filter_pos_ix is not None:
        keep = np.flatnonzero(np.isin(common[0], filter_pos_ix))
        common = (common[0][keep], common[1][keep])
    n_acc = db_snps.shape[1]
    score = np.zeros(n_acc, dtype=np.float64)
    ninfo = np.zeros(n_acc, dtype=np.int64)
    m = len(common[0])
    s_wei = np.asarray(s_wei, dtype=np.float64)
    for j in range(0, m, chunk_size):
        rows = common[0][j:j + chunk_size]
        t_s, t_n = match_gts_accs(s_wei[common[1][j:j + chunk_size]], db_snps[rows, :], skip_db_hets)
        score = score + t_s
        ninfo = ninfo + t_n
    overlap = get_fraction(m, len(s_pos))
    return InbredResult(score, ninfo, overlap, m, common)


# --------------------------------------------------------------------------
# A5/A6 — windows
# --------------------------------------------------------------------------
def genome_chr_ids(ref_chrs):
    """genomes.py:28 — lower-case, 'chr' removed."""
    return np.array([str(c).lower().replace("chr", "") for c in ref_chrs], dtype="str")


def num_windows(chrlen, bin_len):
    """genomes.py:113 — len(range(1, chrlen, bin_len))."""
    return len(range(1, int(chrlen), int(bin_len)))


def window_of(pos, chrlen, bin_len):
    """Window number of a position inside its chromosome, -1 when it is in no window.
    Window k covers [1 + k*b, (k+1)*b] for k < num_windows (genomes.py:113-116)."""
    pos = np.asarray(pos, dtype=np.int64)
    k = (pos - 1) // int(bin_len)
    ok = (pos >= 1) & (k < num_windows(chrlen, bin_len))
    return np.where(ok, k, -1)


class WindowResult(object):
    pass


def window_genotyper(db_snps, db_chrs, db_chr_regions, db_positions, s_chrs, s_pos, s_wei,
                     genome_chrs, genome_chrlen, bin_len, skip_db_hets=False, error_rate=0.02):
    """csmatch.py:64-104 with genomes.py:73-127: per window join + matchGTsAccs on the whole
    window, totals in window order, per-window epilogue (csmatch.py:44-61).

    Returns a WindowResult with per-window arrays (only windows with >=1 matched marker carry
    data; `win_has` marks them) and the totals."""
    db_positions = np.asarray(db_positions)
    s_pos = np.asarray(s_pos)
    s_wei = np.asarray(s_wei, dtype=np.float64)
    g_ids = genome_chr_ids(genome_chrs)
    db_ids = genome_chr_ids(db_chrs)
    s_ids_all = genome_chr_ids(s_chrs)
    n_acc = db_snps.shape[1]
    tot_score = np.zeros(n_acc, dtype=np.float64)
    tot_ninfo = np.zeros(n_acc, dtype=np.int64)
    matched_tar = []
    win_rows = []
    num_mat = 0
    win_index = 1
    winds_chrs = []
    for ci, cid in enumerate(g_ids):
        hit = np.flatnonzero(db_ids == cid)
        if len(hit):
            start, end = int(db_chr_regions[hit[0]][0]), int(db_chr_regions[hit[0]][1])
        else:
            start, end = 0, 0
        g_rows = np.arange(start, end)
        g_win = window_of(db_positions[start:end], genome_chrlen[ci], bin_len)
        s_rows = np.flatnonzero(s_ids_all == cid)
        s_win = window_of(s_pos[s_rows], genome_chrlen[ci], bin_len)
        for k in range(num_windows(genome_chrlen[ci], bin_len)):
            e_g = g_rows[g_win == k]
            e_s = s_rows[s_win == k]
            gp = db_positions[e_g]
            sp = s_pos[e_s]
            acc_ind = e_g[np.isin(gp, sp)]
            tar_ind = e_s[np.isin(sp, gp)]
            num_mat += len(acc_ind)
            if len(acc_ind) > 0:
                sc, ni = match_gts_accs(s_wei[tar_ind], db_snps[acc_ind, :], skip_db_hets)
                tot_score = tot_score + sc
                tot_ninfo = tot_ninfo + ni
                matched_tar.append(tar_ind)
                win_rows.append((win_index, sc, ni))
            winds_chrs.append(cid)
            win_index += 1
    res = WindowResult()
    res.n_windows = win_index - 1
    res.windows = win_rows                    # list of (window_index, score f64[A], ninfo i64[A])
    res.tot_score = tot_score
    res.tot_ninfo = tot_ninfo
    res.num_snps = num_mat
    res.overlap = get_fraction(num_mat, len(s_pos))
    res.matched_tar = np.concatenate(matched_tar) if matched_tar else np.zeros(0, dtype=np.int64)
    res.winds_chrs = np.array(winds_chrs, dtype="str")
    res.error_rate = error_rate
    return res


def window_epilogue(score, ninfo, error_rate=0.02):
    """csmatch.py:44-61 for one window: (L, LR, identical, num_amb, keep_mask)."""
    lik, lr = calculate_likelihoods(score, ninfo)
    ident = test_identity(score, ninfo, error_rate=error_rate)
    with np.errstate(invalid="ignore"):
        amb = lr < LR_THRES
    num_amb = int(amb.sum())
    keep = amb if (1 <= num_amb < len(score)) else np.zeros(len(score), dtype=bool)
    return lik, lr, ident, num_amb, keep


# --------------------------------------------------------------------------
# A7 — simulated F1 pass
# --------------------------------------------------------------------------
def top_hit_accessions(probs, k=10):
    """csmatch.py:113."""
    return np.argsort(-np.asarray(probs))[0:k]


def f1_pair_scores(db_snps, common, s_wei, top_accs):
    """csmatch.py:115-126 — 45 simulated F1s of the top-10 accessions over the whole-genome join."""
    s_wei = np.asarray(s_wei, dtype=np.float64)
    w = s_wei[common[1]]
    pairs, scores, ninfos = [], [], []
    for i, j in itertools.combinations(top_accs, 2):
        g1 = db_snps[common[0], i]
        g2 = db_snps[common[0], j]
        homalt = np.flatnonzero((g1 == 1) & (g2 == 1))
        homref = np.flatnonzero((g1 == 0) & (g2 == 0))
        het = np.flatnonzero((g1 != -1) & (g2 != -1) & (g1 != g2))
        sc = np.sum(w[homalt, 2]) + np.sum(w[homref, 0]) + np.sum(w[het, 1])
        pairs.append((int(i), int(j)))
        scores.append(sc)
        ninfos.append(len(homalt) + len(homref) + len(het))
    return pairs, np.array(scores, dtype=np.float64), np.array(ninfos, dtype=np.int64)


def segregating_rows(db_snps, accs_ix):
    """snp_genotype.py:188-211 + segregting_snps (:378-383): after masking missing calls (< 0 -> nan) and sorting each row,
    t_sum = 1 + #adjacent equal pairs, t_r_sum = #called; rows with t_sum / t_r_sum < 1 and t_r_sum != 0."""
    t = np.array(np.asarray(db_snps)[:, np.asarray(accs_ix)], dtype=float)
    t[t < 0] = np.nan
    t = np.sort(t, axis=1)
    t_r_sum = np.sum(~np.isnan(t), axis=1)
    t_sum = np.nansum(t[:, 1:] == t[:, :-1], axis=1) + 1
    with np.errstate(divide="ignore", invalid="ignore"):
        div = np.where(t_r_sum != 0, t_sum / np.maximum(t_r_sum, 1), np.inf)
    return np.setdiff1d(np.flatnonzero(div < 1), np.flatnonzero(t_r_sum == 0))


# --------------------------------------------------------------------------
# 8(f)-3 — pairsnp and simulate
# --------------------------------------------------------------------------
def pairwise_score(chrs1, pos1, gt1, chrs2, pos2, gt2, name1="1", name2="2", db=None):
    """snpmatch.py:270-309 — pairwiseScore on parsed inputs.  db = (db_chromosome_labels, db_positions) restricts sample 1
    to the database positions first (:276-281).  Returns the stats dict of the reference (without the 'hdf5' path)."""
    chrs1, chrs2 = np.asarray(chrs1, dtype="str"), np.asarray(chrs2, dtype="str")
    pos1, pos2 = np.asarray(pos1), np.asarray(pos2)
    gt1, gt2 = np.asarray(gt1, dtype="str"), np.asarray(gt2, dtype="str")
    if db is not None:
        c1 = get_common_positions(db[0], db[1], chrs1, pos1)
        ci = get_common_positions(chrs1[c1[1]], pos1[c1[1]], chrs2, pos2)
        ci = (c1[1][ci[0]], ci[1])
    else:
        ci = get_common_positions(chrs1, pos1, chrs2, pos2)
    unique_1 = len(chrs1) - len(ci[0])
    unique_2 = len(chrs2) - len(ci[0])
    g1, g2 = normalize_chr_names(chrs1), normalize_chr_names(chrs2)
    stats = {}
    common, scores = [], []
    for i in np.intersect1d(first_appearance_ids(g1), first_appearance_ids(g2)):
        per = np.flatnonzero(g1[ci[0]] == i)
        t_common = len(per)
        t_scores = int(np.sum(gt1[ci[0][per]] == gt2[ci[1][per]]))
        stats[str(i)] = [get_fraction(t_scores, t_common), t_common]
        common.append(t_common)
        scores.append(t_scores)
    stats["matches"] = [get_fraction(int(np.sum(scores)), int(np.sum(common))), int(np.sum(common))]
    stats["unique"] = {name1: [get_fraction(unique_1, len(chrs1)), len(chrs1)], name2: [get_fraction(unique_2, len(chrs2)), len(chrs2)]}
    return stats


def simulate_snps(acc_snp, row_chrs, row_pos, num_snps, err_rate, rng=np.random):
    """simulate.py:10-31 — draws in the reference's order from `rng` (np.random there): marker subset, rows to corrupt,
    replacement calls.  Returns (chr, pos, binary call) of the simulated sample."""
    acc_snp = np.asarray(acc_snp)
    informative = np.flatnonzero(acc_snp >= 0)
    pick = np.sort(rng.choice(np.arange(informative.shape[0]), num_snps, replace=False))
    rows = informative[pick]
    snp = acc_snp[rows].astype(np.int8)
    n_change = int(err_rate * num_snps)
    values = rng.choice(3, n_change)            # simulate.py:26 is an assignment: its right-hand side is drawn first
    change = np.sort(rng.choice(np.arange(num_snps), n_change, replace=False))
    snp[change] = values
    return np.asarray(row_chrs)[rows], np.asarray(row_pos)[rows], snp


def simulate_snps_f1(snps_p1, snps_p2, row_chrs, row_pos, num_snps, err_rate, rm_hets=1, rng=np.random):
    """simulate.py:33-60."""
    p1, p2 = np.asarray(snps_p1), np.asarray(snps_p2)
    common_ix = np.flatnonzero((p1 >= 0) & (p2 >= 0) & (p1 < 2) & (p2 < 2))
    seg = np.flatnonzero(p1[common_ix] != p2[common_ix])
    same = np.setdiff1d(np.arange(len(common_ix)), seg)
    common_snps = np.zeros(len(common_ix), dtype="int8")
    common_snps[seg] = 2
    common_snps[same] = p1[common_ix[same]]
    pick = np.sort(rng.choice(np.arange(len(common_ix)), num_snps, replace=False))
    snp = common_snps[pick].astype(int)
    n_change = int(err_rate * num_snps)
    values = rng.choice(2, n_change)            # simulate.py:52: right-hand side first
    change = np.sort(rng.choice(np.flatnonzero(snp != 2), n_change, replace=False))
    snp[change] = values
    het_ix = np.flatnonzero(snp == 2)
    snp[het_ix] = rng.choice(3, het_ix.shape[0], p=[(1 - rm_hets) / 2, (1 - rm_hets) / 2, rm_hets])
    rows = common_ix[pick]
    return np.asarray(row_chrs)[rows], np.asarray(row_pos)[rows], snp.astype(np.int8)


# --------------------------------------------------------------------------
# 8(f)-4 — genotype_cross: windowed parent matching
# --------------------------------------------------------------------------
def parse_gt(gt):
    """parsers.py:12-35 — GT strings -> 0 / 1 / 2 (het) / -1 (no call); anything else is 0.  Separator from the first element."""
    gt = np.asarray(gt, dtype="str")
    out = np.zeros(len(gt), dtype=np.int8)
    if len(gt) == 0:
        return out
    sep = "|" if "|" in gt[0] else "/"
    out[gt == "1" + sep + "1"] = 1
    out[(gt == "0" + sep + "1") | (gt == "1" + sep + "0")] = 2
    out[gt == "." + sep + "."] = -1
    return out


def get_window_genotype(matched, total, lr_thres, n_marker_thres=5):
    """genotype_cross.py:21-49 — matched = [parent 1, het, parent 2] counts of a window.  Returns 0, 1, 2 or 'NA'."""
    if total < n_marker_thres:
        return "NA"
    assert len(matched) == 3
    if np.array_equal(np.array(matched), np.repeat(0, 3)):
        return "NA"
    lik, lr = calculate_likelihoods(matched, np.repeat(total, 3))
    if len(np.where(lr == 1)[0]) > 1:
        return 1
    high = int(np.nanargmin(lik))
    rest = lr[np.nonzero(lr - 1)]
    lr_next = np.nan if (len(rest) == 0 or np.all(np.isnan(rest))) else np.nanmin(rest)
    if np.isnan(lr_next):
        lr_next = lr_thres
    geno = "NA"
    if high == 0 and lr_next >= lr_thres:
        geno = 0
    elif high == 2 and lr_next >= lr_thres:
        geno = 2
    if high == 1:
        geno = 1
    return geno


def segregating_parent_markers(snps_p1, snps_p2):
    """genotype_cross.py:108 — rows on which both parents are called and differ."""
    p1, p2 = np.asarray(snps_p1), np.asarray(snps_p2)
    return np.flatnonzero((p1 != p2) & (p1 >= 0) & (p2 >= 0))


def genotype_cross_windows(par_chrs, par_pos, par_p1, par_p2, vcf_chrs, vcf_pos, vcf_gt, genome_chrs, genome_chrlen, bin_len, lr_thres):
    """genotype_cross.py:210-241 — per genome window (JSON order) the list [geno per sample] or None when the window holds no
    matched marker, plus the counts.  vcf_gt: GT strings [n, S].  Returns (calls: list over windows, counts: dict w -> int [S,3],
    n_matched per window)."""
    ids = genome_chr_ids(genome_chrs)
    p_ids, v_ids = genome_chr_ids(par_chrs), genome_chr_ids(vcf_chrs)
    par_pos, vcf_pos = np.asarray(par_pos), np.asarray(vcf_pos)
    vcf_gt = np.asarray(vcf_gt, dtype="str")
    calls, counts, n_matched = [], {}, []
    w = 0
    for ci, cid in enumerate(ids):
        p_rows, v_rows = np.flatnonzero(p_ids == cid), np.flatnonzero(v_ids == cid)
        p_win = window_of(par_pos[p_rows], genome_chrlen[ci], bin_len)
        v_win = window_of(vcf_pos[v_rows], genome_chrlen[ci], bin_len)
        for k in range(num_windows(genome_chrlen[ci], bin_len)):
            e_b, e_s = p_rows[p_win == k], v_rows[v_win == k]
            acc_ind = e_b[np.isin(par_pos[e_b], vcf_pos[e_s])]
            tar_ind = e_s[np.isin(vcf_pos[e_s], par_pos[e_b])]
            n_matched.append(len(tar_ind))
            if len(tar_ind) == 0:
                calls.append(None)
            else:
                row, cnt = [], []
                for s_ix in range(vcf_gt.shape[1]):
                    t = parse_gt(vcf_gt[tar_ind, s_ix])
                    m = [int(np.sum(t == par_p1[acc_ind])), int(np.sum(t == 2)), int(np.sum(t == par_p2[acc_ind]))]
                    cnt.append(m)
                    row.append(get_window_genotype(m, len(tar_ind), lr_thres))
                calls.append(row)
                counts[w] = np.array(cnt)
            w += 1
    return calls, counts, n_matched


# --------------------------------------------------------------------------
# data-format helper shared by the tests (not reference behaviour)
# --------------------------------------------------------------------------
def pack_2bit_words(db_snps):
    """Reference packing used by tests to check the device packer: code = int8 & 3
    (0 ref, 1 alt, 2 het, 3 missing); per row one uint64 per 32 accessions, low half = bit 0 of
    the codes, high half = bit 1 (accession g*32+j is bit j); the row is padded to an even number
    of words and padding accessions are missing (3).  See include/snpmatch_b200.h."""
    db = np.asarray(db_snps, dtype=np.int8)
    n, a = db.shape
    words = (a + 31) // 32
    stride = (words + 1) & ~1
    codes = np.full((n, stride * 32), 3, dtype=np.uint8)
    codes[:, :a] = (db & 3).astype(np.uint8)
    lo = np.packbits((codes & 1).reshape(n, stride, 32), axis=2, bitorder="little").view("<u4").reshape(n, stride)
    hi = np.packbits((codes >> 1).reshape(n, stride, 32), axis=2, bitorder="little").view("<u4").reshape(n, stride)
    return np.ascontiguousarray(lo.astype(np.uint64) | (hi.astype(np.uint64) << np.uint64(32)))
