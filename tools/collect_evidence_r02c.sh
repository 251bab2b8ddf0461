#!/bin/bash
# End-of-round refresh of the launch list and the per-kernel metrics (the planner now runs the 4x4 layers as single CTAs);
# same commands as collect_evidence_r02b.sh, outputs gpurun_out/ev5_*.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=74
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ev5_launches.csv \
  python bench.py --steps 1 --warmup 3 --reverse-steps 3 --e2e-steps 0 --gpu-eager 0 > gpurun_out/ev5_launches.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:'k_conv_tc|k_gn_apply|k_groupnorm|k_split_input|k_attention|k_gemv_rows' \
  --launch-skip $N --launch-count $N --csv --log-file gpurun_out/ev5_metrics.csv python tools/profile_forward.py > gpurun_out/ev5_metrics.log 2>&1
