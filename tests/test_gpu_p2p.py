"""The one-shot reduce over peer memory (snpm_batch_ipc_* / snpm_batch_reduce_peers) needs two GPUs and one process
per GPU: scripts/check_p2p_reduce.py is launched under torchrun and compares it with the NCCL reduce-scatter bit for bit.
Skipped on a single-GPU box (the host logic of the sharded layouts is covered by tests/test_sharding_gloo.py on CPU)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_reduce_equals_nccl_reduce_scatter():
    import __graft_entry__ as ge
    ge.build()
    from snpmatch_b200 import lib
    if lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_p2p_reduce.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["all_ranks_ok"], r
    assert r["checks"]["p2p_score_rel"] == 0.0 and r["checks"]["exact_p2p_matches"]


def test_peer_reduce_single_rank_and_errors():
    """world = 1 degenerates to a copy of the own rows; call-order errors are reported, not crashed on."""
    import numpy as np
    import __graft_entry__ as ge
    ge.build()
    from snpmatch_b200 import lib, synth
    panel = synth.small_panel(n_rows=20000, n_acc=300)
    sample = synth.make_sample(panel["positions"], panel["chr_regions"], panel["chrs"], 300, true_acc=7, n_db=3000, n_extra=300, seed=501)
    db = lib.Database(panel["positions"], panel["chr_regions"], 300)
    db.load_int8(panel["snps"])
    b = lib.Batch(db, [0, len(sample["pos"])], sample["chr_ix"], sample["pos"], sample["wei"])
    b.run()
    with pytest.raises(lib.SnpmError):
        b.reduce_peers()                       # nothing opened
    b.epilogue()
    want = b.fetch()
    h = b.ipc_export()
    assert len(h) == 64
    b.ipc_open([h], 0)
    for _ in range(3):                         # the flags step along; the head-of-run wait passes at once
        b.run()
        b.reduce_peers()
        b.epilogue()
        got = b.fetch()
        assert np.array_equal(got["score"], want["score"]) and np.array_equal(got["ninfo"], want["ninfo"])
    b.ipc_close()
    b.close()
    db.close()
