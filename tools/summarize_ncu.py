"""Turn an ncu launch list (CSV of gpu__time_duration.sum per launch) and/or an `--set full`
report into the short text summaries committed under profiles/.

  python tools/summarize_ncu.py launches  gpurun_out/launches.csv          > profiles/rX_launches.md
  python tools/summarize_ncu.py kernel    gpurun_out/prof.ncu-rep          > profiles/rX_kernels.md
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[h + 1:]:
        if len(r) > vi:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            tot[r[ki]] += v
            cnt[r[ki]] += 1
    T = sum(tot.values())
    print(f"# ncu launch list: {path}\n\nper-launch times are cold-cache and serialised: compare SHARES.\n")
    print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:25]:
        print(f"| {100 * v / T:.1f}% | {v / 1e3:.1f} | {cnt[k]} | {v / 1e3 / cnt[k]:.1f} | `{k[:110]}` |")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    print(f"# ncu --set full: {path}\n")
    for r in rows[2:]:
        print(f"## `{r[ni][:100]}`\n\n| metric | unit | value |\n|---|---|---|")
        vals = dict(zip(hdr, r))
        for k in KEYS:
            if k in vals:
                print(f"| {k} | {units[hdr.index(k)]} | {vals[k]} |")
        try:
            rd = float(vals["dram__bytes_read.sum"].replace(",", ""))
            wr = float(vals["dram__bytes_write.sum"].replace(",", ""))
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            tr = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
            print(f"| traffic = dram read + write | byte | {tr:.0f} |")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
