#!/bin/bash
# A/B of library builds on one GPU box: every build/ab/*.so is copied over the in-tree library and the grouped kernel is
# timed on the headline workload (1 GPU) and on rank 0's share of an 8-GPU run (scripts/measure_grouped.py --shard-of 8).
# Usage (under gpurun): bash scripts/ab_variants.sh
cd "$(dirname "$0")/.."
cp snpmatch_b200/libsnpmatch_b200.so /tmp/lib_keep.so
for f in build/ab/*.so; do
    cp "$f" snpmatch_b200/libsnpmatch_b200.so
    touch snpmatch_b200/libsnpmatch_b200.so
    n=$(basename "$f" .so)
    python scripts/measure_grouped.py --samples 64 --chunks 320 --reps 7 > gpurun_out/ab_${n}_n1.json 2> gpurun_out/ab_${n}_n1.err
    python scripts/measure_grouped.py --samples 512 --shard-of 8 --chunks 320 --reps 7 > gpurun_out/ab_${n}_s8.json 2> gpurun_out/ab_${n}_s8.err
    python - "$n" <<'P'
import json, sys
n = sys.argv[1]
for tag in ("n1", "s8"):
    try:
        d = json.load(open("gpurun_out/ab_%s_%s.json" % (n, tag)))["grouped_chunk_320"]
        print(n, tag, "score %.4f combine %.4f total %.4f matches_equal %s ninfo_equal %s score_rel %.1e" % (
            d["score_ms"], d["combine_ms"], d["total_ms"], d.get("matches_equal"), d.get("ninfo_equal"), d.get("score_max_rel", -1)))
    except Exception as e:
        print(n, tag, "FAILED", e)
P
done
cp /tmp/lib_keep.so snpmatch_b200/libsnpmatch_b200.so
