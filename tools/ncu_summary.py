"""Text summary of an .ncu-rep (one kernel): the roofline-relevant raw metrics + the most-stalled SASS lines.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rN_x.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"# {rep}\n# kernel: {name}\n")
    for h, u, v in zip(hdr, units, vals):
        base = h.split(".", 2)[-1] if h.split(".")[0].isupper() else h
        if h in KEYS or any(h.endswith(k) for k in KEYS):
            print(f"{h:100s} {u:16s} {v}")
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    hdr = src[1]
    data = src[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data) or 1
    agg = collections.Counter()
    for r in data:
        for i in stall:
            agg[hdr[i]] += int(r[i])
    print(f"\n# warp-state samples: {tot}; by reason: " + ", ".join(f"{k}={v}" for k, v in agg.most_common(8)))
    print("# most sampled SASS instructions (samples, share, executed, instruction, top stall reasons)")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:30]:
        st = sorted(((int(r[i]), hdr[i]) for i in stall), reverse=True)[:2]
        print(f"{int(r[isamp]):7d} {100 * int(r[isamp]) / tot:5.1f}% {r[iex]:>10s}  {r[isrc].strip()[:72]:72s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")


if __name__ == "__main__":
    main(sys.argv[1])
