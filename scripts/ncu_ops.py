"""Opcode histogram + hottest SASS lines of one kernel in an .ncu-rep:
python scripts/ncu_ops.py rep kernel_regex [regions] [skip=N]   (skip = matching launches to skip)"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
skip = next((a.split("=")[1] for a in sys.argv if a.startswith("skip=")), "0")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
isrc, ins, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = []
for r in rows[hi + 1:]:
    if len(r) <= max(ist, ins, iw, iwi) or not r[ins].isdigit():
        if len(r) > 1 and r[0] == "Kernel Name":
            break
        continue
    data.append((r[isrc].strip(), int(r[ins]), int(r[ist] or 0), int(r[iw] or 0), int(r[iwi] or 0)))
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data) or 1
print("total warp-instructions", tot, "stall samples", tots, "SASS lines", len(data))
op, ops, wv, wvi = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for s, n, st, w, wi in data:
    t = s.split()
    o = t[1] if t[0].startswith("@") else t[0]
    op[o] += n; ops[o] += st; wv[o] += w; wvi[o] += wi
for o, c in op.most_common(28):
    print(f"{o:22s} {c:12d} {100*c/tot:5.1f}%  stalls {100*ops[o]/tots:5.1f}%  smem wavefronts {wv[o]:10d} ideal {wvi[o]:10d}")
print("--- top stall lines")
for i in sorted(range(len(data)), key=lambda i: -data[i][2])[:25]:
    print(f"{i:5d} {data[i][2]:6d} {data[i][1]:10d}  {data[i][0][:90]}")
if len(sys.argv) > 3 and sys.argv[3] == "regions":
    print("--- regions (start line, #lines, executions per line, share of all instructions)")
    i = 0
    while i < len(data):
        j = i
        while j + 1 < len(data) and abs(data[j + 1][1] - data[i][1]) <= 0.02 * max(1, data[i][1]):
            j += 1
        n = j - i + 1
        share = 100 * sum(d[1] for d in data[i:j + 1]) / tot
        if share > 0.4:
            print(f"{i:5d} {n:4d} {data[i][1]:10d} {share:5.1f}%   {data[i][0][:50]} ... {data[j][0][:40]}")
        i = j + 1
