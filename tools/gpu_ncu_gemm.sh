#!/bin/bash
# ncu --set full of the GEMM kernels only (exported to CSV on the box)
mkdir -p gpurun_out
timeout 120 python tools/prof_kernels.py gemm > gpurun_out/r02_plain_gemm.log 2>&1 && \
timeout 400 ncu --clock-control none --set full --import-source on -k regex:tgemm_kernel -c 9 -o gpurun_out/r02_ncu_full_tgemm_v2 python tools/prof_kernels.py gemm > gpurun_out/r02_ncu_b2.log 2>&1
ncu -i gpurun_out/r02_ncu_full_tgemm_v2.ncu-rep --page raw --csv > gpurun_out/full.csv 2>/dev/null
python - <<'PY'
import csv, re
rows = list(csv.reader(open("gpurun_out/full.csv")))
keep = re.compile(r"ID|Kernel Name|Block Size|Grid Size|dram__bytes_read\.sum$|dram__bytes_write\.sum$|gpu__time_duration\.sum|sm__pipe_tensor.*cycles_active\.avg\.pct|gpu__dram_throughput\.avg\.pct|sm__throughput\.avg\.pct|launch__registers_per_thread|sm__warps_active\.avg\.pct|lts__t_sector_hit_rate\.pct|smsp__average_warps_issue_stalled.*|smsp__warp_issue_stalled.*_per_warp_active\.pct")
idx = [i for i, h in enumerate(rows[0]) if keep.search(h)]
with open("gpurun_out/r02_ncu_full_tgemm_v2.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] if i < len(r) else "" for i in idx])
PY
rm -f gpurun_out/r02_ncu_full_tgemm_v2.ncu-rep gpurun_out/full.csv
tail -n 2 gpurun_out/r02_ncu_b2.log; ls -la gpurun_out | grep tgemm_v2
