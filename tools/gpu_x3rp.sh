#!/bin/bash
# bf16x3 row-packed resblock kernel: parity tests, per-resblock timing, A/B of the bf16x3 hot path (bit 1 = without it)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest split-bf16 rp"
timeout 900 python -m pytest tests/test_gpu_round2.py -q -x -k "split_bf16" -s 2>&1 | grep -v Warning | tail -45
echo "== pytest x3 model-level"
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_round2.py -q -x -k "bf16x3 or x3 or config4 or conv_post" 2>&1 | tail -5
echo "== time rp x3"
timeout 600 python tools/time_rb.py --rp 1 --x3 1 --reps 5 2>&1 | tail -8
echo "== A/B bf16x3 hot path"
timeout 600 python tools/ab_hotpath.py --precision bf16x3 1 3 2>&1 | tail -6
