#!/bin/bash
# A/B of library builds on ONE GPU box (box-to-box differences are as large as the effects measured): every build/ab/*.so is
# timed on the coded path's measurement workloads.  Usage (under gpurun): bash scripts/ab_coded.sh [rounds]
cd "$(dirname "$0")/.."
rounds=${1:-2}
for r in $(seq 1 $rounds); do
for f in build/ab/*.so; do
    n=$(basename "$f" .so)
    python scripts/measure_coded.py --lib $f --no-old --no-exact --reps 9 > gpurun_out/ab_${n}_n1_$r.json 2> gpurun_out/ab_${n}.err
    python scripts/measure_coded.py --lib $f --no-old --no-exact --reps 7 --samples 512 --shard-of 8 > gpurun_out/ab_${n}_s8_$r.json 2>> gpurun_out/ab_${n}.err
    python scripts/measure_coded.py --lib $f --no-old --no-exact --reps 5 --accessions 20000 --samples 16 > gpurun_out/ab_${n}_wide_$r.json 2>> gpurun_out/ab_${n}.err
    python scripts/measure_coded.py --lib $f --no-old --no-exact --reps 7 --hard > gpurun_out/ab_${n}_hard_$r.json 2>> gpurun_out/ab_${n}.err
    python - $n $r <<'P'
import json, sys
n, r = sys.argv[1], sys.argv[2]
out = [n, r]
for tag in ("n1", "s8", "wide", "hard"):
    try:
        d = json.load(open("gpurun_out/ab_%s_%s_%s.json" % (n, tag, r)))["coded_chunk_320"]
        out.append("%s %.4f" % (tag, d["score_ms"]))
    except Exception as e:
        out.append("%s FAILED" % tag)
print(" ".join(out))
P
done
done
