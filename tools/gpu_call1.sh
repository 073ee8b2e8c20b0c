#!/bin/bash
# GPU call: full GPU test-suite, bench c2 (with the reference-GPU arm), launch list, one full ncu capture of the attention kernel
mkdir -p gpurun_out
rm -f gpurun_out/teacher_forced.log gpurun_out/logit_err.log
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "bench rc=$?" >> gpurun_out/r2_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-micro --no-reference-gpu --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
PROBE_L=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tree_attn_tc -s 10 -c 1 -f -o gpurun_out/r2_attn_prof python tools/attn_probe.py > gpurun_out/r2_ncu_attn.log 2>&1
SQ_ATTN_TIMING=1 timeout 120 python tools/attn_probe.py > gpurun_out/r2_attn_probe.log 2>&1
tail -3 gpurun_out/r2_t1.log; cat gpurun_out/r2_bench_c2.json | head -c 3000; tail -2 gpurun_out/r2_bench_c2.err; cat gpurun_out/r2_attn_probe.log
