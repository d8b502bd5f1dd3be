#!/usr/bin/env python
"""Turn ncu outputs into the markdown summaries kept under profiles/.

  python scripts/summarize_profile.py launches gpurun_out/x_launches.csv [skip_until_kernel_regex]
  python scripts/summarize_profile.py full gpurun_out/x.ncu-rep
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, libonly=True):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    tot = 0.0
    for row in rows:
        name = row["Kernel Name"]
        if libonly and "gcg::" not in name:
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v
        short = re.sub(r"\(.*", "", name.replace("void ", "").replace("gcg::", ""))
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print("| kernel | launches | total ms | avg ms | share |")
    print("|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.3f | %.1f%% |" % (k, c, t, t / c, 100 * t / tot))
    print("\ntotal libgcg kernel time in the capture: %.2f ms over %d launches (cold-cache, serialised: compare shares)" % (tot, sum(c for c, _ in agg.values())))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("\n### `%s` grid %s" % (r[idx["Kernel Name"]], r[idx.get("Grid Size", 0)]))
        print("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in idx:
                print("| %s | %s | %s |" % (w, r[idx[w]], units[idx[w]]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2])
