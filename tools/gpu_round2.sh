#!/bin/bash
# Round-2 evidence run (one B200): full GPU test suite, bench line, ncu launch lists and --set full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round2.sh
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_gputest_final.log
echo "== bench"
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 600 gpurun_out/r2_bench_final.json
echo "== ncu: decoder pass launch list with DRAM bytes"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_decoder_traffic.csv python tools/one_decode.py > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
echo "== ncu: bench step launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench_step.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-parity-mode --no-sharded --no-full-model > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
echo "== ncu --set full: row-packed resblock kernel (C=16, k=7) and the attention kernel"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:rp_tc -c 1 -o gpurun_out/r2_rp16k7 -f \
    python tools/time_rb.py --rp 1 --only 16,7 --reps 2 > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:relenc_attention_bf16 -c 1 -o gpurun_out/r2_attention_bf16 -f \
    python tools/time_relenc.py --reps 1 > gpurun_out/ncu_d.log 2>&1; tail -2 gpurun_out/ncu_d.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_relenc.csv \
    python tools/time_relenc.py --reps 1 > gpurun_out/ncu_e.log 2>&1; tail -2 gpurun_out/ncu_e.log
