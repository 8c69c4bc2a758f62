#!/bin/bash
# A/B of library builds on one GPU box: every build/ab/*.so is copied over the in-tree library and bench.py is run on it.
# Usage (under gpurun): bash scripts/ab_variants.sh [bench args...]
cd "$(dirname "$0")/.."
cp snpmatch_b200/libsnpmatch_b200.so /tmp/lib_keep.so
for f in build/ab/*.so; do
    cp "$f" snpmatch_b200/libsnpmatch_b200.so
    n=$(basename "$f" .so)
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err
    python - "$n" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/ab_%s.json" % n))
    print(n, "value %.3e step %.4f score %.4f join %.4f frac %.3f e2e %.3e called_kernel %.4f" % (
        d["value"], d["ms_per_step"], d["stages_ms"]["score_ms"], d["stages_ms"]["join_ms"], d["roofline"]["frac"], d["e2e"]["value"],
        d["called_genotypes"]["roofline"]["kernel_ms"]))
except Exception as e:
    print(n, "FAILED", e)
P
done
cp /tmp/lib_keep.so snpmatch_b200/libsnpmatch_b200.so
