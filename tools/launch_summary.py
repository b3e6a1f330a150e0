"""Summarise one steady-state step of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_summary.py gpurun_out/launches.csv [--all] [--diff other.csv]
"""
import csv, re, sys, collections

def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1000 if u in ("nsecond", "ns") else v * 1000 if u in ("msecond", "ms") else v
        n = re.sub(r"void |\(anonymous namespace\)::|mocha::|<unnamed>::", "", x["Kernel Name"])
        n = re.sub(r"\(.*$", "", n)
        rows.append((n, v, x.get("Grid Size", "")))
    idx = [i for i, r in enumerate(rows) if "post_frame" in r[0]]
    return rows[idx[-2] + 1: idx[-1] + 1]

step = load(sys.argv[1])
print(f"step launches {len(step)} total us {sum(v for _, v, _ in step):.1f}")
if "--diff" in sys.argv:
    other = load(sys.argv[sys.argv.index("--diff") + 1])
    for i, ((n, v, g), (n2, v2, g2)) in enumerate(zip(step, other)):
        flag = " <<<" if abs(v - v2) > 3 else ""
        print(f"{i:3d} {v:7.1f} {v2:7.1f} {v - v2:+7.1f} {g:>14s} {n[:60]}{flag}")
elif "--all" in sys.argv:
    for i, (n, v, g) in enumerate(step):
        print(f"{i:3d} {v:7.1f} {g:>14s} {n[:70]}")
else:
    agg = collections.defaultdict(lambda: [0.0, 0])
    for n, v, _ in step:
        agg[n][0] += v; agg[n][1] += 1
    tot = sum(v for _, v, _ in step)
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{v:9.1f} us {100 * v / tot:5.1f}%  x{c:3d}  {n[:80]}")
