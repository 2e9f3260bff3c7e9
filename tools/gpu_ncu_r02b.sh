#!/bin/bash
# final round-2 ncu evidence: one `ncu --set full` pass over the kernels changed late in the round (exported to CSV on
# the box; the .ncu-rep is deleted because gpurun_out is merged back only below 64 MiB)
mkdir -p gpurun_out
timeout 200 python tools/prof_kernels.py r02b > gpurun_out/r02b_plain.log 2>&1 || { tail -5 gpurun_out/r02b_plain.log; exit 1; }
timeout 1200 ncu --clock-control none --set full --import-source on -k regex:'tgemm_kernel|attention_|layernorm_|select_fused' -c 24 \
  -o gpurun_out/r02b python tools/prof_kernels.py r02b > gpurun_out/r02b_ncu.log 2>&1
ncu -i gpurun_out/r02b.ncu-rep --page raw --csv > gpurun_out/full.csv 2>/dev/null
python - <<'PY'
import csv, re
rows = list(csv.reader(open("gpurun_out/full.csv")))
keep = re.compile(r"ID|Kernel Name|Block Size|Grid Size|dram__bytes_read\.sum$|dram__bytes_write\.sum$|gpu__time_duration\.sum|sm__pipe_tensor.*cycles_active\.avg\.pct|gpu__dram_throughput\.avg\.pct|sm__throughput\.avg\.pct|launch__registers_per_thread|sm__warps_active\.avg\.pct|lts__t_sector_hit_rate\.pct|smsp__average_warps_issue_stalled.*|smsp__warp_issue_stalled.*_per_warp_active\.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|sm__inst_executed_pipe_xu|smsp__inst_executed\.sum$")
idx = [i for i, h in enumerate(rows[0]) if keep.search(h)]
with open("gpurun_out/r02_ncu_full_final.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] if i < len(r) else "" for i in idx])
PY
rm -f gpurun_out/r02b.ncu-rep gpurun_out/full.csv
tail -n 3 gpurun_out/r02b_ncu.log; ls -la gpurun_out | grep "r02_ncu_full_final"
