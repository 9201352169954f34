"""Group the SASS lines of an .ncu-rep by their executed count (= which loop they sit in) and total the samples.

    python tools/ncu_phases.py gpurun_out/x.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isamp, iex, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
g = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), collections.Counter()])
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    ex = int(r[iex])
    e = g[ex]
    e[0] += 1
    e[1] += int(r[isamp])
    e[2] += ex
    for i in stall:
        e[3][hdr[i]] += int(r[i])
    e[4][r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]] += 1
tot = sum(e[1] for e in g.values())
totex = sum(e[2] for e in g.values())
print(f"# {rep}: {tot} samples, {totex} warp instructions")
print("# executed/line  lines  samples  share  warp-instr share   top stalls | top opcodes")
for ex, e in sorted(g.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{ex:10d} {e[0]:6d} {e[1]:7d} {100 * e[1] / tot:5.1f}% {100 * e[2] / totex:5.1f}%   "
          + " ".join(f"{k[6:]}={v}" for k, v in e[3].most_common(3)) + " | " + " ".join(f"{k}={v}" for k, v in e[4].most_common(6)))
