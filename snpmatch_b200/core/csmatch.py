"""
`snpmatch cross` — host side of the windowed B200 matching path.

Mirrors `snpmatch/core/csmatch.py`: `CrossIdentifier` (csmatch.py:19-186) with `window_genotyper`
(:64-104), `get_window_data` (:44-61), `match_insilico_f1s` (:106-129), `cross_interpreter`
(:131-186) and `potatoCrossIdentifier` (:193-200).  The per-window join, scoring, likelihoods,
identity calls and the simulated-F1 reductions run on the GPU (csrc/windows.cuh, csrc/f1.cuh);
this file turns the device arrays into the reference's `windowscore.txt`, `scores.txt` and JSON
files.
"""
import itertools
import json
import logging

import numpy as np
import pandas as pd

from .. import lib
from . import genomes
from . import parsers
from . import snp_genotype
from . import snpmatch

log = logging.getLogger(__name__)
chunk_size = 1000

WINDOW_COLUMNS = ["acc", "snps_match", "snps_info", "score", "likelihood", "identical", "num_amb", "window_index"]


def _np_str(a):
    """Text NumPy gives a float column inside np.column_stack with a string column (csmatch.py:50)."""
    return np.asarray(a).astype("U32")


class CrossIdentifier(object):

    def __init__(self, inputs, g, genome_id, binLen, output_id="cross.identifier", run_identifier=True,
                 identity_error_rate=0.02, skip_db_hets=False):
        self.g = g
        assert type(inputs) is parsers.ParseInputs, "provide a parsers class"
        inputs.filter_chr_names()
        self.inputs = inputs
        self.genome = genomes.Genome(genome_id)
        self.binLen = binLen
        self.output_id = output_id
        self.error_rate = identity_error_rate
        self._skip_db_hets = skip_db_hets
        self._batch = None
        if run_identifier:
            self.cross_identifier()

    def cross_identifier(self):
        try:
            res = self.window_genotyper(self.output_id + '.windowscore.txt')
            res.print_json_output(self.output_id + ".scores.txt.matches.json")
            snpmatch.getHeterozygosity(self.inputs.gt[res.matchedTarInd], self.output_id + ".scores.txt.matches.json")
            with open(self.output_id + ".scores.txt.matches.json") as json_out:
                self.cross_identfier_json = json.load(json_out)
            self.result = self.match_insilico_f1s(res, self.output_id + '.scores.txt')
            self.cross_interpreter(self.output_id + ".matches.json")
        finally:
            self._close_batch()

    def _close_batch(self):
        if self._batch is not None:
            self._batch.close()
            self._batch = None

    @staticmethod
    def get_window_data(bin_inds, AccList, ScoreList, NumInfoSites, error_rate=0.02):
        """Rows of one window (csmatch.py:44-61) from host arrays — API-compatible entry point; the
        workflow itself gets the same quantities for all windows at once from the device."""
        AccList = np.asarray(AccList)
        ScoreList = np.asarray(ScoreList, dtype=np.float64)
        NumInfoSites = np.asarray(NumInfoSites)
        prob, lik, lrt = lib.calculate_likelihoods(ScoreList, NumInfoSites)
        identity = snpmatch.np_test_identity(ScoreList, NumInfoSites, error_rate=error_rate)
        with np.errstate(invalid="ignore"):
            amb = np.flatnonzero(lrt < snpmatch.lr_thres)
        return _window_frame(bin_inds, AccList, ScoreList, NumInfoSites, prob, lik, identity, len(amb), amb, len(AccList))

    def window_genotyper(self, out_file, mask_acc_ix=None):
        """All windows in one device pass (csmatch.py:64-104)."""
        g = self.g
        num_lines = len(g.accessions)
        if mask_acc_ix is not None:
            assert type(mask_acc_ix) is np.ndarray, "please provide numpy array of acc indices to be masked"
            keep = np.setdiff1d(np.arange(num_lines), mask_acc_ix)
        else:
            keep = np.arange(num_lines)
        binLen = int(self.binLen)
        win_count, win_off, n_windows, winds_chrs = self.genome.window_layout(g.g.chrs, binLen)
        s_ids = genomes.genome_style_ids(self.inputs.chrs)
        uniq = np.unique(s_ids)
        assert len(uniq) <= len(self.genome.chrs_ids), "Please change default --genome option"
        assert len(np.intersect1d(uniq, self.genome.chrs_ids)) > 0, "Please change default --genome option"
        order, cid, pos = g.prepare_markers(self.inputs.chrs, self.inputs.pos, style="genome")
        wei = np.ascontiguousarray(np.asarray(self.inputs.wei, dtype=np.float64)[order])
        # identity table: a window cannot match more markers than the sample has in it
        n_max = int(np.bincount(np.maximum(cid, 0)).max()) if len(cid) else 0
        if len(pos):
            wkey = cid.astype(np.int64) * (1 << 32) + (np.maximum(pos.astype(np.int64), 1) - 1) // binLen
            n_max = int(np.unique(wkey, return_counts=True)[1].max())
        kmax = snpmatch.identity_kmax_table(n_max, self.error_rate)
        self._close_batch()
        batch = g.db.scratch_batch([0, len(pos)], cid, pos, wei)
        self._batch = batch
        self._order = order
        self._join_style_same = np.array_equal(
            g.prepare_markers(self.inputs.chrs, self.inputs.pos, style="join")[1], cid)
        batch.run_windows(self._skip_db_hets, binLen, win_count, win_off, n_windows, kmax, snpmatch.lr_thres)
        batch.epilogue()
        tot = batch.fetch()
        self.timings = batch.timings()
        accs = g.accessions
        frames = []
        masked = mask_acc_ix is not None
        if masked:                                       # likelihoods over the kept accessions only: from the full window arrays
            w = batch.fetch_windows()
            for wi in np.flatnonzero(w["nrows"] > 0):
                sc, ni = w["score"][wi][keep], w["ninfo"][wi][keep]
                frames.append(self.get_window_data(wi + 1, accs[keep], sc, ni, self.error_rate))
        else:                                            # the surviving rows of all windows, compacted on the device
            w = batch.fetch_window_rows()
            frames.append(_window_rows_frame(w, accs))
        frames = [f for f in frames if len(f)]
        self.windows_data = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame(columns=WINDOW_COLUMNS)
        NumMatSNPs = int(tot["m"][0])
        overlap = snpmatch.get_fraction(NumMatSNPs, len(self.inputs.pos))
        result = snpmatch.GenotyperOutput(accs[keep], tot["score"][0][keep], tot["ninfo"][0][keep], overlap, NumMatSNPs,
                                          self.inputs.dp)
        if not masked:
            result._attach_fused(tot["prob"][0], tot["L"][0], tot["LR"][0])
        result.matchedTarInd = order[w["matched_s_idx"]]
        result.winds_chrs = winds_chrs
        if out_file is not None:
            self.windows_data.to_csv(out_file, sep="\t", index=False)
            return result
        return [self.windows_data, result]

    def match_insilico_f1s(self, snpmatch_result, out_file):
        """Simulated F1s of the ten most probable accessions (csmatch.py:106-129)."""
        assert type(snpmatch_result) is snpmatch.GenotyperOutput, "Please provide GenotyperOutput class as input"
        if not hasattr(snpmatch_result, 'probabilies'):
            snpmatch_result.get_probabilities()
        log.info("simulating F1s for top 10 accessions")
        TopHitAccs = np.argsort(-snpmatch_result.probabilies)[0:10]
        g = self.g
        batch = self._batch
        own = False
        if batch is None or not getattr(self, "_join_style_same", False):
            order, cid, pos = g.prepare_markers(self.inputs.chrs, self.inputs.pos)
            wei = np.ascontiguousarray(np.asarray(self.inputs.wei, dtype=np.float64)[order])
            batch = lib.Batch(g.db, [0, len(pos)], cid, pos, wei)
            batch.run(False)
            own = True
        try:
            if len(TopHitAccs) >= 2:
                f_score, f_ninfo = batch.f1_pairs(TopHitAccs)
            else:
                f_score, f_ninfo = np.zeros(0), np.zeros(0, dtype=np.int64)
        finally:
            if own:
                batch.close()
        names = [g.accessions[i] + "x" + g.accessions[j] for i, j in itertools.combinations(TopHitAccs, 2)]
        snpmatch_result.scores = np.append(snpmatch_result.scores, f_score)
        snpmatch_result.ninfo = np.append(snpmatch_result.ninfo, f_ninfo)
        snpmatch_result.accs = np.append(snpmatch_result.accs, np.array(names, dtype="str"))
        if out_file is not None:
            snpmatch_result.print_out_table(out_file)
        return snpmatch_result

    def cross_interpreter(self, out_file):
        """Verdict on the window table (csmatch.py:131-186); writes only for interpretation case >= 3."""
        assert 'cross_identfier_json' in dir(self), "run cross identifier first!"
        assert 'windows_data' in dir(self), "run window genotyper first!"
        log.info("running cross interpreter!")
        js = self.cross_identfier_json
        if js['interpretation']['case'] < 3:
            return
        wd = self.windows_data
        acc = wd["acc"].to_numpy().astype("str")
        widx = wd["window_index"].to_numpy().astype(int)
        ident = wd["identical"].to_numpy().astype(float)
        namb = wd["num_amb"].to_numpy().astype(int)
        windows = np.unique(widx)
        win_ident = np.array([ident[widx == k].max() for k in windows]) if len(windows) else np.zeros(0)
        identical_wind = np.flatnonzero(win_ident == 1)     # positions in the sorted window list, as in the reference
        num_winds = len(windows)
        js['identical_windows'] = [snpmatch.get_fraction(len(identical_wind), num_winds), num_winds]
        homo_wind = np.intersect1d(widx[namb < 20], identical_wind)
        in_homo = np.isin(widx, homo_wind)
        h_names, h_counts = np.unique(acc[in_homo], return_counts=True)
        js['matches'] = [(str(h_names[i]), int(h_counts[i])) for i in np.argsort(-h_counts)]
        res = self.result
        topMatch = np.argsort(res.likelis)[0]
        is_f1_row = ~np.isin(res.accs, self.g.accessions)
        if is_f1_row[topMatch]:
            mother, father = res.accs[topMatch].split("x")[0], res.accs[topMatch].split("x")[1]
            js['interpretation'] = {'text': "Sample may be a F1! or a contamination!", 'case': 5}
            js['parents'] = {'mother': [mother, 1], 'father': [father, 1]}
            js['genotype_windows'] = {'chr_bins': None, 'coordinates': {'x': None, 'y': None}}
        else:
            c_names, c_counts = np.unique(acc[namb == 1], return_counts=True)
            if len(c_names) > 0:
                top2 = np.argsort(-c_counts)[0:2]
                parents = c_names[top2].astype("str")
                counts = c_counts[top2].astype("int")
                xdict = np.array(windows, dtype="int")
                ydict = np.repeat("NA", len(xdict)).astype("S25")
                if len(parents) == 1:
                    js['interpretation'] = {'text': "Sample may be a F2! but only one parent found!", 'case': 6}
                    js['parents'] = {'mother': [parents[0], counts[0]], 'father': ["NA", "NA"]}
                    ydict[np.isin(xdict, widx[(acc == parents[0]) & in_homo])] = parents[0]
                    chr_bins = None
                else:
                    js['interpretation'] = {'text': "Sample may be a F2!", 'case': 6}
                    js['parents'] = {'mother': [parents[0], counts[0]], 'father': [parents[1], counts[1]]}
                    n_names, n_counts = np.unique(res.winds_chrs, return_counts=True)
                    chr_bins = dict((n_names[i], n_counts[i]) for i in range(len(n_names)))
                    ydict[np.isin(xdict, widx[(acc == parents[0]) & in_homo])] = parents[0]
                    ydict[np.isin(xdict, widx[(acc == parents[1]) & in_homo])] = parents[1]
                js['genotype_windows'] = {'chr_bins': chr_bins, 'coordinates': {'x': xdict.tolist(), 'y': ydict.tolist()}}
            else:
                js['interpretation'] = {'case': 7, 'text': "Sample may just be contamination!"}
                js['genotype_windows'] = {'chr_bins': None, 'coordinates': {'x': None, 'y': None}}
                js['parents'] = {'mother': [None, 0], 'father': [None, 1]}
        with open(out_file, "w") as out_stats:
            out_stats.write(json.dumps(js, sort_keys=True, indent=4, default=convert_int64))


def _window_frame(bin_inds, accs, score, ninfo, prob, lik, identity, num_amb, amb_rows, num_lines):
    """DataFrame rows of one window: only accessions with LR < lr_thres, and only when at least one
    but not all accessions pass (csmatch.py:57-60).  `score` and `likelihood` are text, as produced by
    the reference's np.column_stack with the accession names (csmatch.py:50)."""
    if not (1 <= num_amb < num_lines):
        return pd.DataFrame(columns=WINDOW_COLUMNS)
    k = np.asarray(amb_rows)
    sc = np.asarray(score, dtype=np.float64)[k]
    frame = pd.DataFrame({
        "acc": np.asarray(accs)[k].astype(str),
        "snps_match": np.array([int(float(s)) for s in _np_str(sc)], dtype=int),
        "snps_info": np.asarray(ninfo)[k].astype(int),
        "score": _np_str(np.asarray(prob, dtype=np.float64)[k]),
        "likelihood": _np_str(np.asarray(lik, dtype=np.float64)[k]),
        "identical": np.asarray(identity)[k].astype(float),
    })
    frame["num_amb"] = int(num_amb)
    frame["window_index"] = int(bin_inds)
    return frame[WINDOW_COLUMNS]


def _window_rows_frame(w, accs):
    """All rows of windowscore.txt at once from the device-compacted rows (Batch.fetch_window_rows): the same columns and
    the same text for `score` / `likelihood` as _window_frame builds window by window."""
    counts = np.diff(w["row_off"]).astype(np.int64)
    if counts.sum() == 0:
        return pd.DataFrame(columns=WINDOW_COLUMNS)
    sc = w["score"]
    ni = w["ninfo"]
    with np.errstate(invalid="ignore", divide="ignore"):
        prob = np.where(ni > 0, sc / ni, np.nan)
    frame = pd.DataFrame({
        "acc": np.asarray(accs)[w["acc"]].astype(str),
        "snps_match": np.array([int(float(x)) for x in _np_str(sc)], dtype=int),
        "snps_info": ni.astype(int),
        "score": _np_str(prob),
        "likelihood": _np_str(w["L"]),
        "identical": w["identical"].astype(float),
        "num_amb": np.repeat(w["num_amb"].astype(int), counts),
        "window_index": np.repeat(np.arange(1, len(counts) + 1, dtype=int), counts),
    })
    return frame[WINDOW_COLUMNS]


def convert_int64(o):
    """json `default` hook of the reference (csmatch.py:188-191): NumPy integers become ints; anything
    else it is asked about (the byte strings of `ydict`) falls through to None, i.e. JSON null."""
    if isinstance(o, np.integer):
        return int(o)


def potatoCrossIdentifier(args):
    inputs = parsers.ParseInputs(inFile=args['inFile'], logDebug=args['logDebug'])
    log.info("loading genotype files!")
    g = snp_genotype.Genotype(args['hdf5File'], args['hdf5accFile'])
    log.info("running cross identifier!")
    CrossIdentifier(inputs, g, args['genome'], args['binLen'], args['outFile'], run_identifier=True,
                    skip_db_hets=args['skip_db_hets'])
    log.info("finished!")
