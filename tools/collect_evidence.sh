#!/bin/bash
# Round evidence: run on the GPU box through gpurun; outputs land in gpurun_out/ev_* and are summarised into profiles/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"
timeout 600 python bench.py > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/ev_bench_ref.json 2>> gpurun_out/ev_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ev_launches.csv \
  python bench.py --steps 1 --warmup 3 --reverse-steps 3 --e2e-steps 0 > gpurun_out/ev_launches.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:'k_conv_tc|k_gn_apply|k_groupnorm|k_split_input|k_attention|k_gemv_rows' \
  --launch-skip 104 --launch-count 104 --csv --log-file gpurun_out/ev_metrics.csv python tools/profile_forward.py > gpurun_out/ev_metrics.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sas_vec|k_stable_A_vec|k_reverse_step_fast|k_lim_step_vec' \
  --launch-skip 6 --launch-count 6 -o gpurun_out/ev_prof_stream -f python tools/profile_stream.py > gpurun_out/ev_prof_stream.log 2>&1
for skip in 55 62 101; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc' --launch-skip $skip --launch-count 1 \
    -o gpurun_out/ev_prof_conv_$skip -f python tools/profile_forward.py > gpurun_out/ev_prof_conv_$skip.log 2>&1
done
timeout 900 python tools/bench_configs.py > gpurun_out/ev_configs.jsonl 2> gpurun_out/ev_configs.err
timeout 300 python tools/bench_stream.py > gpurun_out/ev_stream.jsonl 2>&1
timeout 300 python tools/profile_ops.py > gpurun_out/ev_ops.txt 2>&1
nvidia-smi > gpurun_out/ev_smi.txt
