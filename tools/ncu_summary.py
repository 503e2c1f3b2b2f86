"""Summarise ncu outputs brought back by gpurun into small text files under profiles/.

  python tools/ncu_summary.py launches <launches.csv>            -> per-kernel launch count / total / share
  python tools/ncu_summary.py full <report.ncu-rep> [regex]      -> the roofline-relevant raw metrics per launch
"""
import csv
import collections
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_read\.sum|dram__bytes_write\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|dram__cycles_active\.avg|"
    r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|sm__pipe_tensor_subpipe.*pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_tensor.*\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|"
    r"launch__shared_mem_per_block_dynamic|launch__occupancy_limit.*|lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct|"
    r"l1tex__data_bank_conflicts_pipe_lsu\.sum|smsp__cycles_active\.avg|sm__cycles_elapsed\.max|"
    r"smsp__inst_executed\.sum|sm__inst_executed_pipe_(fma|alu|fmaheavy|uniform)\.sum|"
    r"smsp__average_warp.*issue_stalled.*_per_warp_active\.pct|smsp__warp_issue_stalled.*)$")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        d = per.setdefault(name, [0, 0.0, r[7], r[8]])
        d[0] += 1
        d[1] += float(r[14])
    tot = sum(v[1] for v in per.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total (ncu-serialised, cold cache: compare shares)")
    print(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}  block grid")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:10.1f} {100 * v[1] / tot:6.1f}%  {v[2]} {v[3]}")


def full(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        if pat and not re.search(pat, rec.get("Kernel Name", "")):
            continue
        print(f"## {rec.get('Kernel Name', '')[:110]}  id={rec.get('ID')}")
        for h, u in zip(hdr, units):
            if KEEP.match(h) and rec[h] not in ("", "n/a"):
                print(f"  {h:85s} {rec[h]:>18s} {u}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
