"""Turns ncu exports into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches <launches.csv>            per-kernel device time and share
  python tools/summarize_ncu.py full <report.ncu-rep>               key metrics per profiled launch
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg, order = None, {}, []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            key = (d["Kernel Name"].split("(")[0][:90], d["Grid Size"])
            if key not in agg:
                agg[key] = [0, 0.0]
                order.append(key)
            agg[key][0] += 1
            agg[key][1] += float(d["Metric Value"].replace(",", ""))
    total = sum(v[1] for v in agg.values())
    print(f"{'avg us':>9} {'launches':>8} {'share':>7}  kernel / grid")
    for k in order:
        n, t = agg[k]
        print(f"{t / n / 1e3:9.1f} {n:8d} {100 * t / total:6.1f}%  {k[0]} {k[1]}")
    print(f"total profiled device time {total / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("---", r[idx["Kernel Name"]][:100])
        for k in KEYS:
            if k in idx:
                print(f"   {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
