#!/bin/bash
# Round-2 evidence run (one B200): full GPU test suite, bench line, ncu launch lists and --set full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round2.sh
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_gputest_final.log
echo "== bench"
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 600 gpurun_out/r2_bench_final.json
echo "== flow timing"
timeout 100 python tools/time_flow.py 2>&1 | grep "flow reverse" | tee gpurun_out/r2_time_flow.log
echo "== full model timing"
timeout 200 python tools/time_full_model.py 2>&1 | tail -3 | tee gpurun_out/r2_time_full_model.log
echo "== ncu: decoder pass launch list with DRAM bytes"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_decoder_traffic.csv python tools/one_decode.py > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
echo "== ncu: bf16x3 decoder pass launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_decoder_bf16x3.csv \
    python tools/one_decode.py 16 1000 bf16x3 > gpurun_out/ncu_a3.log 2>&1; tail -1 gpurun_out/ncu_a3.log
echo "== ncu: bench step launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench_step.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-parity-mode --no-sharded --no-full-model > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
if [ -n "$VSG_SET_FULL" ]; then
echo "== ncu --set full: row-packed resblock kernel (C=32, k=11), the last conv2 of a C=64 resblock and conv_post on tcgen05"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:rp_tc -c 1 -o gpurun_out/r2_rp32k11 -f \
    python tools/time_rb.py --rp 1 --only 32,11 --reps 2 > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:rp_tc -c 1 -o gpurun_out/r2_rp32k11_x3 -f \
    python tools/time_rb.py --rp 1 --x3 1 --only 32,11 --reps 2 > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:conv_tc_kernelILb1ELi7E -c 1 -o gpurun_out/r2_c64_final -f \
    python tools/one_decode.py > gpurun_out/ncu_d.log 2>&1; tail -2 gpurun_out/ncu_d.log
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:conv_tc_kernelILb1ELi8E -c 1 -o gpurun_out/r2_conv_post_tc -f \
    python tools/one_decode.py > gpurun_out/ncu_e.log 2>&1; tail -2 gpurun_out/ncu_e.log
fi
