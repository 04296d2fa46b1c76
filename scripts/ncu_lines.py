"""Per-source-line samples/instructions from an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_lines.py rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = None; H = None; data = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": H = r; continue
    if H is None or len(r) < len(H) - 2: continue
    if r[2] != "-": continue            # SASS rows carry an address; source rows have "-"
    try:
        ns = int(r[H.index("# Samples")]); ie = int(r[H.index("Instructions Executed")])
    except ValueError:
        continue
    data.append((ns, ie, cur, int(r[0]), r[1]))
ts = sum(d[0] for d in data) or 1; ti = sum(d[1] for d in data) or 1
print("total samples", ts, "total warp instructions", ti)
byfile = collections.Counter()
for d in data: byfile[d[2]] += d[0]
print(dict(byfile))
if len(sys.argv) > 3:      # ranges: name:lo-hi,...
    for spec in sys.argv[3].split(","):
        name, rg = spec.split(":"); lo, hi = map(int, rg.split("-"))
        s = sum(d[0] for d in data if d[2].startswith("svd_small") and lo <= d[3] <= hi)
        i = sum(d[1] for d in data if d[2].startswith("svd_small") and lo <= d[3] <= hi)
        print(f"{name:12s} lines {lo}-{hi}: {100*s/ts:5.1f}% samples {100*i/ti:5.1f}% instructions")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100*d[0]/ts:5.1f}% samp {100*d[1]/ti:5.1f}% inst {d[2]}:{d[3]}: {d[4].strip()[:100]}")
