"""Top stall sites of one kernel from an ncu report (SASS-level source page). usage: ncu_hot.py rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first kernel only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
h = rows[0]
si, ss = h.index("Source"), h.index("# Samples")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[ss] or 0) for r in rows[1:])
print(f"total samples {tot}")
agg = {}
for r in rows[1:]:
    for i in stall_cols:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
idx = sorted(range(1, len(rows)), key=lambda i: -int(rows[i][ss] or 0))[:n]
for i in sorted(idx):
    r = rows[i]
    top = sorted(((int(r[j] or 0), h[j]) for j in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[ss]):6d} {100*int(r[ss])/tot:5.1f}%  {r[si][:90]:90s} {top}")
