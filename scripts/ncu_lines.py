"""Aggregate an ncu source export (`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`) by CUDA source line:
python scripts/ncu_lines.py export.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sections, cur, i = [], None, 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        cur = {'file': r[1], 'rows': []}
        sections.append(cur)
    elif r and r[0] == "Line No":
        cur['hdr'] = r
    elif cur is not None and r and r[0] != "Function Name":
        cur['rows'].append(r)
    i += 1
for sec in sections:
    hdr, ix = sec['hdr'], {}
    for k, h in enumerate(hdr):
        ix.setdefault(h, k)
    agg = {}
    for r in sec['rows']:
        if len(r) < len(hdr) or not r[0].strip().isdigit():
            continue
        try:
            smp, ins = int(r[ix['# Samples']] or 0), int(r[ix['Instructions Executed']] or 0)
        except ValueError:
            continue
        a = agg.setdefault(int(r[0]), [0, 0, r[1][:100]])
        a[0] += smp
        a[1] += ins
    tot, toti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    if tot < 50:
        continue
    print(sec['file'], 'samples', tot, 'warp instructions', toti)
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(ln, a[0], '%.1f%%' % (100 * a[0] / max(tot, 1)), a[1], '%.1f%%' % (100 * a[1] / max(toti, 1)), a[2])
