"""Top stall sites of one kernel launch in an .ncu-rep: python scripts/ncu_hot.py rep kernel_regex [launch_index] [threshold_pct]"""
import csv, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.8
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]; data = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    if len(r) == len(hdr): data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
print(rows[start - 1][1][:100])
S = "# Samples" if "# Samples" in ci else "Warp Stall Sampling (All Samples)"
tot = sum(int(r[ci[S]]) for r in data)
print("total samples", tot, "instructions", len(data))
for k, r in enumerate(data):
    s = int(r[ci[S]])
    if s > tot * thr / 100:
        print(f"{k:5d} {100*s/tot:5.1f}%  exec={r[ci['Instructions Executed']]:>8s}  {r[ci['Source']][:90]}")
b = collections.Counter()
for k, r in enumerate(data):
    b[k // 200] += int(r[ci[S]])
print("by 200-instr bucket:", {k * 200: round(100 * v / tot, 1) for k, v in sorted(b.items())})
