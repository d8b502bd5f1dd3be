#!/bin/bash
# round-2 last call: ncu --set full of the FINAL streaming SpMM (cursor, 128 non-zeros per span): DRAM traffic per launch
mkdir -p gpurun_out
BENCH="python bench.py --workload twitter-world --steps 1 --warmup 3 --no-cpu-baseline --no-parity"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:spmm_stream_kernel" -s 6 -c 2 -f -o /tmp/fin_spmm $BENCH > gpurun_out/fin_ncu_spmm.log 2>&1
echo "ncu spmm rc=$?"
ncu -i /tmp/fin_spmm.ncu-rep --page raw --csv > gpurun_out/fin_ncu_spmm_stream.raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/fin_ncu_spmm_stream.raw.csv")))
h = rows[0]
for r in rows[2:]:
    g = lambda k: r[h.index(k)] if k in h else "?"
    print(g("Kernel Name")[:60], "| time", g("gpu__time_duration.sum"), "| read", g("dram__bytes_read.sum"), "write", g("dram__bytes_write.sum"),
          "| L2 hit", g("lts__t_sector_hit_rate.pct"), "| dram %", g("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
          "| lts %", g("lts__throughput.avg.pct_of_peak_sustained_elapsed"), "| grid", g("launch__grid_size"))
print([u for u in rows[1]][h.index("dram__bytes_read.sum")] if "dram__bytes_read.sum" in h else "")
PY
