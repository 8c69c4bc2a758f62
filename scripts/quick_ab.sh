#!/bin/bash
# Development helper (under gpurun): the counting kernel of the coded path on the four measurement workloads.
# Usage: bash scripts/quick_ab.sh TAG
cd "$(dirname "$0")/.."
t=$1
show() { python - "$1" <<'P'
import json, sys
d = json.load(open(sys.argv[1]))["coded_chunk_320"]
print(sys.argv[1].split("/")[-1], "score %.4f group %.4f join %.4f combine %.4f" % (d["score_ms"], d["group_ms"], d["join_ms"], d["combine_ms"]),
      "eq", d.get("matches_equal"), d.get("ninfo_equal"), "rel %.1e" % d.get("score_max_rel", -1), "flagged", d["guard_flagged_samples"])
P
}
python scripts/measure_coded.py --no-old --reps 9 > gpurun_out/${t}_n1.json 2> gpurun_out/${t}_n1.err; show gpurun_out/${t}_n1.json
python scripts/measure_coded.py --no-old --no-exact --reps 9 > gpurun_out/${t}_n1b.json 2> gpurun_out/${t}_n1b.err; show gpurun_out/${t}_n1b.json
python scripts/measure_coded.py --no-old --reps 7 --samples 512 --shard-of 8 > gpurun_out/${t}_s8.json 2> gpurun_out/${t}_s8.err; show gpurun_out/${t}_s8.json
python scripts/measure_coded.py --no-old --no-exact --reps 7 --hard > gpurun_out/${t}_hard.json 2> gpurun_out/${t}_hard.err; show gpurun_out/${t}_hard.json
python scripts/measure_coded.py --no-old --no-exact --reps 5 --accessions 20000 --samples 16 > gpurun_out/${t}_wide.json 2> gpurun_out/${t}_wide.err; show gpurun_out/${t}_wide.json
