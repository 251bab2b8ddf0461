#!/bin/bash
# Round-2 (final build) evidence: run on the GPU box through gpurun; outputs land in gpurun_out/ev4_* and are summarised into
# profiles/r02_*.  One forward of the default engine = 3 k_gemv_rows + 71 ops (54 convolutions, 15 k_gn_apply, split, attention).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=74
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"
# launch list of the bench command (graph kernel nodes are profiled one by one)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ev4_launches.csv \
  python bench.py --steps 1 --warmup 3 --reverse-steps 3 --e2e-steps 0 --gpu-eager 0 > gpurun_out/ev4_launches.log 2>&1
# every kernel of one UNet forward with DRAM / L2 / TMA bytes and tensor-pipe activity
timeout 900 ncu --metrics $M --clock-control none -k regex:'k_conv_tc|k_gn_apply|k_groupnorm|k_split_input|k_attention|k_gemv_rows' \
  --launch-skip $N --launch-count $N --csv --log-file gpurun_out/ev4_metrics.csv python tools/profile_forward.py > gpurun_out/ev4_metrics.log 2>&1
# full captures: the GroupNorm-in-the-epilogue kernel (16x16, 256 -> 256), a 32x32 128 -> 128 layer, the dx-stacked final conv
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'k_conv_tc<\(int\)256, \(int\)64, \(int\)2, \(int\)1, \(bool\)0, \(bool\)0, \(int\)1>' --launch-skip 9 --launch-count 1 \
  -o gpurun_out/ev4_prof_gne16 -f python tools/profile_forward.py > gpurun_out/ev4_prof_gne16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'k_conv_tc<\(int\)128, \(int\)64, \(int\)2, \(int\)2, \(bool\)0, \(bool\)0, \(int\)0>' --launch-skip 14 --launch-count 1 \
  -o gpurun_out/ev4_prof_conv32 -f python tools/profile_forward.py > gpurun_out/ev4_prof_conv32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'k_conv_tc<\(int\)48' --launch-skip 1 --launch-count 1 \
  -o gpurun_out/ev4_prof_final -f python tools/profile_forward.py > gpurun_out/ev4_prof_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_fill6' --launch-skip 1 --launch-count 1 \
  -o gpurun_out/ev4_prof_fill6 -f python tools/profile_stream.py > gpurun_out/ev4_prof_fill6.log 2>&1
# sanitizers on the smoke path and on the GroupNorm-in-the-epilogue convolutions
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python __graft_entry__.py --smoke > gpurun_out/ev4_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/ev4_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python __graft_entry__.py --smoke > gpurun_out/ev4_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/ev4_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_unet_ops.py -q -x -k "in_the_epilogue and (B40_16x16 or B11_4x4 or B5_8x8 or B33_16x16)" > gpurun_out/ev4_memcheck_gne.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/ev4_memcheck_gne.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_unet_ops.py -q -x -k "in_the_epilogue and (B40_16x16 or B11_4x4 or B5_8x8)" > gpurun_out/ev4_racecheck_gne.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/ev4_racecheck_gne.log
nvidia-smi > gpurun_out/ev4_smi.txt
