#!/usr/bin/env python
"""Two or more ranks (torchrun): the one-shot peer reduce (sharding.p2p_reduce_scatter) against the NCCL reduce-scatter
(sharding.reduce_scatter_batch) and against a single-GPU run of the same samples; prints one JSON line on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_p2p_reduce.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    import bench
    from snpmatch_b200 import lib, sharding, synth
    from snpmatch_b200.core import snp_genotype
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    host_pg = dist.new_group(backend="gloo")
    stream = torch.cuda.Stream(device=dev)
    n_rows, n_acc, S = 2_000_000, 1135, 16 * world
    positions, regions = synth.panel_positions(n_rows)
    r0, r1 = sharding.shard_rows(n_rows, world, rank)
    g = snp_genotype.Genotype.synthetic(n_rows, n_acc, row_range=(r0, r1), device=local)
    g.db.set_stream(stream.cuda_stream)
    samples = bench.make_samples(positions, regions, n_acc, S, 9000)
    offs = np.concatenate([[0], np.cumsum([len(s["pos"]) for s in samples])]).astype(np.int64)
    chrom = np.concatenate([s["chr_ix"] for s in samples]).astype(np.int32)
    pos = np.concatenate([s["pos"] for s in samples]).astype(np.int32)
    wei = np.concatenate([s["wei"] for s in samples])
    gs = lib.group_markers(offs, chrom, pos, wei)
    S_loc = S // world
    res = {}
    with torch.cuda.stream(stream):
        for how in ("nccl", "p2p", "p2p_again", "exact_p2p"):
            b = lib.Batch(g.db, offs, chrom, pos, wei)
            b.set_result_range(rank * S_loc, S_loc)
            grouped = how != "exact_p2p"
            if grouped:
                b.upload_grouped(gs)
            times = []
            for it in range(4):
                if how != "nccl":
                    sharding.p2p_before_run(b, dist, rank, world, host_pg)
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                b.run(kernel_mode=lib.KERNEL_GROUPED if grouped else lib.KERNEL_FP64)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if how == "nccl":
                    sharding.reduce_scatter_batch(b, dist, dev, rank, world)
                else:
                    sharding.p2p_reduce_scatter(b)
                e1.record(stream)
                b.epilogue()
                r = b.fetch()
                times.append(e0.elapsed_time(e1))
            res[how] = {k: v.copy() for k, v in r.items()}
            res[how]["reduce_ms"] = float(np.median(times[1:]))
            if how != "nccl":
                torch.cuda.synchronize()
                dist.barrier()
                b.ipc_close()
            b.close()
    ok = {}
    for how in ("p2p", "p2p_again"):
        ok[how] = all(np.array_equal(res[how][k], res["nccl"][k], equal_nan=True) for k in ("matches", "ninfo", "m", "prob"))
        ok[how + "_score_rel"] = float(np.max(np.abs(res[how]["score"] - res["nccl"]["score"]) / np.maximum(res["nccl"]["score"], 1.0)))
    ok["exact_p2p_matches"] = bool(np.array_equal(res["exact_p2p"]["matches"], res["nccl"]["matches"]) and np.array_equal(res["exact_p2p"]["ninfo"], res["nccl"]["ninfo"]))
    flags = torch.tensor([int(all(v for k, v in ok.items() if isinstance(v, bool)))], device=dev)
    dist.all_reduce(flags)
    if rank == 0:
        print(json.dumps({"world": world, "all_ranks_ok": int(flags[0]) == world, "checks": ok, "nccl_reduce_scatter_ms": res["nccl"]["reduce_ms"],
                          "p2p_reduce_ms": res["p2p"]["reduce_ms"], "samples": S, "payload_bytes_per_rank": int(S * (3 * n_acc + 2) * 8)}), flush=True)
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
