"""Print the interesting fields of a bench.py JSON line (development helper)."""
import json
import sys


def show(k, v, ind=0, width=160):
    if isinstance(v, dict):
        print(" " * ind + k + ":")
        for a, b in v.items():
            show(a, b, ind + 2, width)
    else:
        print(" " * ind + k + ": " + str(v)[:width])


lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
d = json.loads(lines[-1])
keys = sys.argv[2:] or ["value", "ms_per_step", "e2e", "stages_ms", "roofline", "sub_cross", "sub_wide", "sub_api", "sub_a9", "sub_called", "cpu_baseline", "clocks"]
for k in keys:
    if k in d:
        show(k, d[k])
