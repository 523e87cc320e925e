"""Turns ncu CSV exports (gpurun_out/) into the small summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/launches_X.md
  python profiles/summarize.py kernel   gpurun_out/prof_X.ncu-rep  > profiles/kernel_X.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, {}
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        a = agg.setdefault(d["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "at::" not in k and "at_cuda" not in k)
    print(f"| kernel | launches | total ms | share of all | share of this repo's kernels |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        mine = "at::" not in k and "at_cuda" not in k
        if a[1] / tot < 0.0005 and not mine:
            continue
        name = re.sub(r"\(.*", "", k).replace("<unnamed>::", "").replace("void ", "")[:70]
        print(f"| `{name}` | {a[0]} | {a[1]:.3f} | {a[1] / tot:.3f} | {a[1] / ours:.3f} |" if mine else
              f"| `{name}` (torch, synthetic-input generation) | {a[0]} | {a[1]:.3f} | {a[1] / tot:.3f} | |")
    print(f"\ntotal {tot:.2f} ms over all launches, {ours:.2f} ms in this repo's kernels "
          "(ncu per-launch times: cold-cache and serialised; compare SHARES)")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = {n: (u[i], v[i]) for i, n in enumerate(h)}
        print(f"### {d['Kernel Name'][1][:90]}\n\n| metric | value | unit |\n|---|---|---|")
        for m in KEY_METRICS:
            if m in d:
                print(f"| {m} | {d[m][1]} | {d[m][0]} |")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h, data = rows[1], rows[2:]
    isrc, iex, ith = h.index("Source"), h.index("Instructions Executed"), h.index("Avg. Threads Executed")
    tot = sum(int(r[iex]) for r in data)
    ops = collections.Counter()
    for r in data:
        t = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
        ops[t.split()[0].split(".")[0]] += int(r[iex])
    print(f"\nexecuted warp instructions: {tot:,}; by opcode (share):")
    print(", ".join(f"{k} {v / tot:.3f}" for k, v in ops.most_common(16)))
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_")]
    st = collections.Counter()
    for r in data:
        for i in stall_cols:
            try:
                st[h[i]] += int(r[i])
            except ValueError:
                pass
    ts = sum(st.values()) or 1
    print("\nwarp-stall samples (share): " + ", ".join(f"{k[6:]} {v / ts:.3f}" for k, v in st.most_common(8)))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
