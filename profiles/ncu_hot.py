#!/usr/bin/env python
"""Top warp-stall sampling sites of one kernel from a .ncu-rep captured with --import-source on (SASS view):
    python profiles/ncu_hot.py gpurun_out/x.ncu-rep [N]
prints, for the N most-sampled instructions, samples, executed count and the dominant stall reasons."""
import csv, subprocess, sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("# %s: %d instructions, %d samples" % (path, len(body), tot))
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print("# by reason:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    n = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print("%5d %5.1f%% exec %8s  %-58s %s" % (i, 100.0 * n / max(tot, 1), r[col["Instructions Executed"]], r[col["Source"]].strip()[:58],
                                             " ".join("%s:%d" % (w, c) for c, w in why if c)))
