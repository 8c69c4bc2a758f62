import sys, time, json, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge; ge.build()
from snpmatch_b200 import lib, synth
from snpmatch_b200.core import snp_genotype
n_rows, n_acc = 10_700_000, 1135
g = snp_genotype.Genotype.synthetic(n_rows, n_acc)
rng = np.random.default_rng(0)
for S, K in ((4096, 4000), (4096, 20000), (1024, 100000)):
    rows = np.sort(rng.choice(n_rows, size=K, replace=False))
    codes = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(S, K), p=[0.6, 0.28, 0.02, 0.1])
    best = 1e9
    for _ in range(3):
        r = g.db.score_shared_panel(rows, codes, likelihoods=False)
        best = min(best, r["gemm_ms"])
    macs = (2 * ((S + 63) // 64) * 64) * 1152 * 4 * ((K + 31) // 32 * 32)
    comps = S * K * n_acc
    print(json.dumps({"config": "configs[3]: %d samples x %d shared markers vs 1135 x 10.7M (one-hot int8 GEMM, tcgen05)" % (S, K),
                      "gemm_ms": best, "int8_TOPS": 2 * macs / (best * 1e-3) / 1e12, "frac_of_4500_TOPS": 2 * macs / (best * 1e-3) / 4.5e15,
                      "comparisons_per_s": comps / (best * 1e-3)}), flush=True)
