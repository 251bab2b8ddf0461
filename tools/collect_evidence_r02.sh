#!/bin/bash
# Round-2 evidence: run on the GPU box through gpurun; outputs land in gpurun_out/ev2_* and are summarised into profiles/r02_*.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"
# launch list of the bench command (graph kernel nodes are profiled one by one)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ev2_launches.csv \
  python bench.py --steps 1 --warmup 3 --reverse-steps 3 --e2e-steps 0 --gpu-eager 0 > gpurun_out/ev2_launches.log 2>&1
# every kernel of one UNet forward with DRAM / L2 / TMA bytes and tensor-pipe activity
timeout 900 ncu --metrics $M --clock-control none -k regex:'k_conv_tc|k_gn_apply|k_groupnorm|k_split_input|k_attention|k_gemv_rows' \
  --launch-skip 107 --launch-count 107 --csv --log-file gpurun_out/ev2_metrics.csv python tools/profile_forward.py > gpurun_out/ev2_metrics.log 2>&1
# full capture of the dominant kernel (largest 16x16 N=256 layer and a 32x32 N=128 layer)
for skip in 55 62; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc' --launch-skip $skip --launch-count 1 \
    -o gpurun_out/ev2_prof_conv_$skip -f python tools/profile_forward.py > gpurun_out/ev2_prof_conv_$skip.log 2>&1
done
# sanitizers on the smoke path (noise, Sigma scan, UNet chain through the engine, captured loop) and on the POST convolution
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python __graft_entry__.py --smoke > gpurun_out/ev2_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/ev2_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python __graft_entry__.py --smoke > gpurun_out/ev2_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/ev2_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_unet_ops.py -q -x -k "producer_side and (B3_32x32 or B5_8x8 or B11_4x4 or B5_16x16)" > gpurun_out/ev2_memcheck_post.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/ev2_memcheck_post.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_unet_ops.py -q -x -k "producer_side and (B3_32x32 or B5_8x8 or B11_4x4)" > gpurun_out/ev2_racecheck_post.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/ev2_racecheck_post.log
timeout 300 python bench.py --impl reference > gpurun_out/ev2_bench_ref.json 2> gpurun_out/ev2_bench_ref.err
nvidia-smi > gpurun_out/ev2_smi.txt
