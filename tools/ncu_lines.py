"""Per CUDA source line totals of an .ncu-rep captured with --import-source on (kernel compiled with -lineinfo).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys


def main(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr = rows[hdr_i]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    iline = hdr.index("Line") if "Line" in hdr else None
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr) and r[isamp].isdigit()]
    tot = sum(int(r[isamp]) for r in data) or 1
    totex = sum(int(r[iex]) for r in data if r[iex].isdigit()) or 1
    print(f"# {rep}: {tot} samples, {totex} warp instructions")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
        st = sorted(((int(r[i]), hdr[i]) for i in stall), reverse=True)[:2]
        ln = r[iline] if iline is not None else ""
        print(f"{int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  ex {100 * int(r[iex] or 0) / totex:5.1f}%  {ln:>4s} {r[isrc].strip()[:90]:90s} "
              f"{st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
