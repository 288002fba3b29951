"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one steady-state training step."""
import csv, re, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 6
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = []
for row in csv.DictReader(lines):
    try:
        rows.append((row['Kernel Name'], float(row['Metric Value'].replace(',', ''))))
    except Exception:
        pass
names = [r[0] for r in rows]
marker = sys.argv[4] if len(sys.argv) > 4 else 'nll_loss_forward'
idx = [i for i, n in enumerate(names) if marker in n]
seg = rows[idx[which]:idx[which + 1]]
def short(n):
    n = n.replace('<unnamed>::', '').replace('void ', '')
    return re.sub(r'\(.*', '', n)[:80]
agg = {}
for n, v in seg:
    k = short(n); agg.setdefault(k, [0, 0]); agg[k][0] += v; agg[k][1] += 1
tot = sum(v for _, v in seg)
print("one step = %d launches, %.0f us serialised" % (len(seg), tot / 1000))
print("| kernel | launches | us | share |\n|---|---|---|---|")
for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print("| `%s` | %d | %.1f | %.1f%% |" % (n, c, v / 1000, 100 * v / tot))
if len(sys.argv) > 3:
    print("\nsequence of our kernels (us):")
    for n, v in seg:
        if short(n).startswith('k_'):
            print("%8.1f  %s" % (v / 1000, short(n)))
