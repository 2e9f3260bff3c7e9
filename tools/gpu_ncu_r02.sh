#!/bin/bash
# round-2 ncu evidence.  Reports are exported to CSV on the box and deleted (gpurun_out is merged back only below 64 MiB).
mkdir -p gpurun_out
NCU="ncu --clock-control none"
RAW='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|gpu__dram_throughput|sm__throughput.avg.pct|launch__registers_per_thread|launch__grid_size|sm__warps_active.avg.pct|lts__t_sector_hit_rate|smsp__cycles_active.avg'
timeout 200 python tools/ncu_step.py flickr8k > gpurun_out/r02_plain_step.log 2>&1 && \
timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_ncu_launches_step_flickr8k.csv python tools/ncu_step.py flickr8k > gpurun_out/r02_ncu_a.log 2>&1
timeout 200 python tools/ncu_step.py vitb16 parity 64 > gpurun_out/r02_plain_step4.log 2>&1 && \
timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_ncu_launches_step_vitb16_b64.csv python tools/ncu_step.py vitb16 parity 64 > gpurun_out/r02_ncu_a4.log 2>&1
export_csv() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw_full.csv 2>/dev/null; python - "$1" <<'PY'
import csv, re, sys
name = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/{name}_raw_full.csv")))
keep = re.compile(r"ID|Kernel Name|Block Size|Grid Size|dram__bytes_read\.sum$|dram__bytes_write\.sum$|gpu__time_duration\.sum|sm__pipe_tensor.*cycles_active\.avg\.pct|sm__inst_executed_pipe_tensor.*\.sum$|gpu__dram_throughput\.avg\.pct|sm__throughput\.avg\.pct|launch__registers_per_thread|sm__warps_active\.avg\.pct|lts__t_sector_hit_rate\.pct|dram__throughput\.avg\.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|smsp__warp_issue_stalled.*_per_warp_active\.pct")
hdr = rows[0]
idx = [i for i, h in enumerate(hdr) if keep.search(h)]
with open(f"gpurun_out/{name}.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] if i < len(r) else "" for i in idx])
PY
rm -f gpurun_out/$1.ncu-rep gpurun_out/$1_raw_full.csv; }
timeout 120 python tools/prof_kernels.py gemm > gpurun_out/r02_plain_gemm.log 2>&1 && \
timeout 400 $NCU --set full --import-source on -k regex:tgemm_kernel -c 9 -o gpurun_out/r02_ncu_full_tgemm python tools/prof_kernels.py gemm > gpurun_out/r02_ncu_b.log 2>&1
export_csv r02_ncu_full_tgemm
timeout 120 python tools/prof_kernels.py codec > gpurun_out/r02_plain_codec.log 2>&1 && \
timeout 400 $NCU --set full --import-source on -k regex:'codec_batched|absmax_scale|split_flat|split_scaled_cluster|select_pass|filter_kernel' -c 14 -o gpurun_out/r02_ncu_full_codec python tools/prof_kernels.py codec > gpurun_out/r02_ncu_c.log 2>&1
export_csv r02_ncu_full_codec
timeout 120 python tools/prof_kernels.py loss > gpurun_out/r02_plain_loss.log 2>&1 && \
timeout 300 $NCU --set full --import-source on -k regex:'rowkth|row_stats|grad_kernel|finalize_kernel' -c 7 -o gpurun_out/r02_ncu_full_loss python tools/prof_kernels.py loss > gpurun_out/r02_ncu_d.log 2>&1
export_csv r02_ncu_full_loss
ls -la gpurun_out | grep r02_ | head -30; du -sh gpurun_out
for f in a a4 b c d; do tail -n 2 gpurun_out/r02_ncu_$f.log; done
