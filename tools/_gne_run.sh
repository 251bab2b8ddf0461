set -x
timeout 600 python -m pytest tests/test_gpu_unet_ops.py -x -q -k "epilogue or producer_side" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q -k "full_width_golden or forward_golden" 2>&1 | tail -8
