"""Top stall sites of one kernel in an .ncu-rep (SASS view): python tools/ncu_hot.py <rep> <kernel substring> [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may match: take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
si, ai, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ai] or 0) for r in body)
print("kernel", rows[hdr_i - 1][1][:90], "instructions", len(body), "samples", tot)
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall totals:", ", ".join("%s %.1f%%" % (k, 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ai] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    top = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% x%-8s %-70s %s" % (i, 100 * int(r[ai] or 0) / max(tot, 1), r[ii], r[si].strip()[:70],
                                          " ".join("%s:%d" % (h[6:], v) for v, h in top if v)))
