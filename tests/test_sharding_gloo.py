"""world_size-2 test of the multi-GPU host logic on CPU (gloo): SNP-row shards, local chromosome
ranges, one all-reduce of the packed per-sample totals, epilogue on the reduced totals.  The per-shard
compute stand-in is the CPU oracle (the CUDA kernels need a GPU; their sharded run is in bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import snpmatch_oracle as orc
    from snpmatch_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    golden = os.path.join(ROOT, "tests", "golden")
    p = dict(np.load(os.path.join(golden, "small_panel.npz")))
    s = dict(np.load(os.path.join(golden, "sample_inbred.npz")))
    n_rows, n_acc = p["snps"].shape
    r0, r1 = sharding.shard_rows(n_rows, world, rank)
    regions = sharding.local_regions(p["chr_regions"], r0, r1)
    res = orc.genotyper(p["snps"][r0:r1], p["chrs"], regions, p["positions"][r0:r1], s["chrs"], s["pos"], s["wei"])
    buf = sharding.pack_reduce_rows(res.score_f64, res.ninfo, res.num_snps)
    t = torch.from_numpy(buf)
    dist.all_reduce(t)
    score, ninfo, m = sharding.unpack_reduce_rows(t.numpy(), n_acc)
    # global pair indices of this shard
    # grouped counting path: the shard's totals travel as (F, ninfo, m, I); finalised after the reduce
    rows, tar = res.common
    F, I, ni = sharding.grouped_partials(s["wei"][tar], p["snps"][r0:r1][rows])
    gt = torch.from_numpy(sharding.pack_grouped_rows(F, ni, len(rows), I))
    dist.all_reduce(gt)
    g_score, g_matches, g_ninfo, g_m, g_guard = sharding.finalize_grouped(gt.numpy(), n_acc)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), score=score[0], ninfo=ninfo[0], m=m,
             db_idx=res.common[0] + r0, s_idx=res.common[1], regions=regions, g_score=g_score[0], g_matches=g_matches[0],
             g_ninfo=g_ninfo[0], g_m=g_m, g_guard=g_guard[0])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_row_sharded_allreduce_matches_single_run(tmp_path, world):
    import torch.multiprocessing as mp
    from oracle import snpmatch_oracle as orc
    from snpmatch_b200 import sharding
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = load_golden("inbred_pl.npz")
    outs = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(world)]
    # every rank holds the same reduced totals
    for o in outs[1:]:
        assert np.array_equal(o["score"], outs[0]["score"]) and np.array_equal(o["ninfo"], outs[0]["ninfo"])
    o = outs[0]
    assert int(o["m"][0]) == int(g["num_snps"])
    assert np.array_equal(o["ninfo"], g["ninfo"])                       # integers: exact in any order
    single = orc.genotyper(*[load_golden("small_panel.npz")[k] for k in ("snps", "chrs", "chr_regions", "positions")],
                           *[load_golden("sample_inbred.npz")[k] for k in ("chrs", "pos", "wei")])
    np.testing.assert_allclose(o["score"], single.score_f64, rtol=1e-13)  # fp64: another summation order
    guard = sharding.truncation_guard(o["score"])
    assert np.array_equal(o["score"].astype(np.int64)[~guard], g["scores"][~guard])
    assert guard.sum() <= 1
    # likelihoods on the reduced totals agree with the single run
    lik, lr = orc.calculate_likelihoods(o["score"].astype(np.int64), o["ninfo"])
    np.testing.assert_allclose(lik, g["likelis"], rtol=1e-12, equal_nan=True)
    # grouped path: integers exact wherever the guard is silent, scores to 1e-12, identical on every rank
    for x in outs[1:]:
        assert np.array_equal(x["g_matches"], o["g_matches"]) and np.array_equal(x["g_score"], o["g_score"])
    assert int(o["g_m"][0]) == int(g["num_snps"]) and np.array_equal(o["g_ninfo"], g["ninfo"])
    ok = ~o["g_guard"]
    assert ok.sum() >= len(ok) - 1
    assert np.array_equal(o["g_matches"][ok], g["scores"][ok])
    assert np.array_equal(o["g_score"].astype(np.int64), o["g_matches"])         # the stored score truncates to matches
    np.testing.assert_allclose(o["g_score"], single.score_f64, rtol=1e-12)
    # the shards' pairs concatenate to the whole join
    db_idx = np.concatenate([x["db_idx"] for x in outs])
    s_idx = np.concatenate([x["s_idx"] for x in outs])
    assert np.array_equal(db_idx, g["common_db"]) and np.array_equal(s_idx, g["common_s"])


def test_shard_geometry():
    from snpmatch_b200 import sharding
    reg = np.array([[0, 100], [100, 250], [250, 400]])
    cuts = [sharding.shard_rows(400, 3, r) for r in range(3)]
    assert cuts == [(0, 133), (133, 266), (266, 400)]
    assert sharding.local_regions(reg, 133, 266).tolist() == [[0, 0], [0, 117], [117, 133]]
    assert sharding.local_regions(reg, 0, 133).tolist() == [[0, 100], [100, 133], [133, 133]]
    total = sum(int((sharding.local_regions(reg, a, b)[:, 1] - sharding.local_regions(reg, a, b)[:, 0]).sum()) for a, b in cuts)
    assert total == 400
    buf = sharding.pack_reduce_rows(np.array([[1.5, 2.5]]), np.array([[3, 4]]), [7])
    assert buf.shape == (1, 6) and buf[0].tolist() == [1.5, 2.5, 3.0, 4.0, 7.0, 0.0]
    sc, ni, m = sharding.unpack_reduce_rows(buf * 2, 2)
    assert sc.tolist() == [[3.0, 5.0]] and ni.tolist() == [[6, 8]] and m.tolist() == [14]
    g = sharding.truncation_guard(np.array([4719.0, 4718.999999999999, 12.5, 3.0000000000000004, 0.0]))
    assert g.tolist() == [False, True, False, True, False]
    # grouped totals: I + F; F just below / at / above an integer k >= 1 is flagged, F < 1 never is
    F = np.array([[0.25, 1e-13, 0.9999999999999999, 2.0000000000000004, 3.5, 0.0]])
    I = np.array([[10, 4000, 7, 7, 0, 12]])
    buf = sharding.pack_grouped_rows(F, I + 5, [100], I)
    sc, ma, ni, m, gd = sharding.finalize_grouped(buf, 6)
    assert ma.tolist() == [[10, 4000, 7, 9, 3, 12]] and gd.tolist() == [[False, False, True, True, False, False]]
    assert sc.astype(np.int64).tolist() == ma.tolist() and ni.tolist() == [[15, 4005, 12, 12, 5, 17]] and m.tolist() == [100]
    w = np.array([[1.0, 0.5, 0.25], [0.125, 1.0, 1.0], [1.0, 0.0, 0.75]])
    codes = np.array([[0, 1, 2, -1], [2, 1, 0, 1], [1, 1, -1, 0]], dtype=np.int8)
    Fp, Ip, Np = sharding.grouped_partials(w, codes)
    assert Ip.tolist() == [2, 1, 0, 2] and Np.tolist() == [3, 3, 2, 2]
    assert Fp.tolist() == [0.75, 1.0, 0.625, 0.0]


def test_marker_slices_cover_the_join_exactly():
    from oracle import snpmatch_oracle as orc
    from snpmatch_b200 import sharding, synth
    pos, regions = synth.panel_positions(50000)
    s = synth.make_sample_fast(pos, regions, 50, 3, n_db=4000, n_extra=500, seed=11)
    rows = s["rows"]
    for world in (1, 2, 3, 8):
        seen = np.zeros(len(rows), dtype=int)
        for rank in range(world):
            r0, r1 = sharding.shard_rows(50000, world, rank)
            i0, i1 = sharding.shard_marker_range(s["chr_ix"], s["pos"], regions, pos, r0, r1)
            in_shard = (rows >= r0) & (rows < r1)
            assert in_shard[:i0].sum() == 0 and in_shard[i1:].sum() == 0      # nothing matchable is left out
            seen[i0:i1] += 1
        assert np.all(seen[rows >= 0] >= 1)
        assert seen.sum() <= len(rows) + 2 * world                             # slices overlap by boundary markers at most
    assert sharding.shard_marker_range(s["chr_ix"], s["pos"], regions, pos, 10, 10) == (0, 0)


def _cross_worker(rank, world, port, out_dir):
    """`cross` on a sharded panel (the exchange step of sharding.run_windows_sharded) with the oracle as the compute stand-in."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import snpmatch_oracle as orc
    from snpmatch_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_acc, bin_len = 20, 3000000
    panel = synth.small_panel(n_rows=3000, n_acc=n_acc)
    s = synth.make_sample(panel["positions"], panel["chr_regions"], panel["chrs"], n_acc, true_acc=3, n_db=900, n_extra=40, seed=77)
    chrlen = synth.TAIR10_CHRLEN
    per_chr = [orc.num_windows(l, bin_len) for l in chrlen]
    base = np.concatenate([[0], np.cumsum(per_chr)])
    W = int(base[-1])
    r0, r1 = sharding.shard_rows(len(panel["positions"]), world, rank)
    reg = sharding.local_regions(panel["chr_regions"], r0, r1)
    i0, i1 = sharding.shard_marker_range(s["chr_ix"], s["pos"], panel["chr_regions"], panel["positions"], r0, r1)
    part = orc.window_genotyper(panel["snps"][r0:r1], panel["chrs"], reg, panel["positions"][r0:r1], s["chrs"][i0:i1], s["pos"][i0:i1],
                                s["wei"][i0:i1], panel["chrs"], chrlen, bin_len)
    score, ninfo, nrows = np.zeros((W, n_acc)), np.zeros((W, n_acc), dtype=np.int64), np.zeros(W, dtype=np.int64)
    for widx, sc, ni in part.windows:
        score[widx - 1], ninfo[widx - 1] = sc, ni
    tar = part.matched_tar                                   # rows per window of this shard, from its matched markers
    win = base[s["chr_ix"][i0:i1][tar]] + (s["pos"][i0:i1][tar] - 1) // bin_len
    np.add.at(nrows, win.astype(np.int64), 1)
    buf = torch.from_numpy(sharding.window_partials_host(score, ninfo, nrows))
    dist.all_reduce(buf)
    sc, ni, nr = sharding.unpack_window_partials(buf.numpy(), W, n_acc)
    np.savez(os.path.join(out_dir, "cross%d.npz" % rank), score=sc, ninfo=ni, nrows=nr)
    dist.destroy_process_group()


def test_cross_window_partials_sum_over_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    from oracle import snpmatch_oracle as orc
    from snpmatch_b200 import synth
    world = 2
    mp.spawn(_cross_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(str(tmp_path / ("cross%d.npz" % r))) for r in range(world)]
    for k in ("score", "ninfo", "nrows"):
        assert np.array_equal(outs[0][k], outs[1][k])       # every rank holds the same sums
    panel = synth.small_panel(n_rows=3000, n_acc=20)
    s = synth.make_sample(panel["positions"], panel["chr_regions"], panel["chrs"], 20, true_acc=3, n_db=900, n_extra=40, seed=77)
    full = orc.window_genotyper(panel["snps"], panel["chrs"], panel["chr_regions"], panel["positions"], s["chrs"], s["pos"], s["wei"],
                                panel["chrs"], synth.TAIR10_CHRLEN, 3000000)
    o = outs[0]
    assert int(o["nrows"].sum()) == full.num_snps and int((o["nrows"] > 0).sum()) == len(full.windows)
    for widx, fs, fn in full.windows:
        assert np.array_equal(o["ninfo"][widx - 1], fn)
        np.testing.assert_allclose(o["score"][widx - 1], fs, rtol=1e-12, atol=0)
