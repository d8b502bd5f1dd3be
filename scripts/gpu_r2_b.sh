#!/bin/bash
# round 2, call B: ncu --set full of the register-gather kernel and streaming variants (same operands).
# Reports stay in /tmp on the box; only the CSV exports (and one report) travel back (gpurun_out <= 64 MiB).
mkdir -p gpurun_out
for cfg in "-1 rows" "2 rows" "3 rows" "7 rows"; do
  set -- $cfg
  tag="v$1_$(echo $2 | tr ':' '_')"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:spmm_(vec|stream)_kernel" -s 2 -c 1 -f -o /tmp/b_ncu_$tag \
     python scripts/spmm_stream_one.py --variant $1 --schedule $2 --span 256 --reps 4 > gpurun_out/b_ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"
  ncu -i /tmp/b_ncu_$tag.ncu-rep --page raw --csv > gpurun_out/b_ncu_$tag.raw.csv 2>/dev/null
  ncu -i /tmp/b_ncu_$tag.ncu-rep --page details --csv > gpurun_out/b_ncu_$tag.details.csv 2>/dev/null
done
cp /tmp/b_ncu_v7_rows.ncu-rep gpurun_out/
ls -la gpurun_out/
