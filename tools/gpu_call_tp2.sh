#!/bin/bash
# 2-GPU call: TP parity tests (one-shot + two-shot all-reduce, tiny + 70B-shaped), bench c2 at N=2 with tp_parity
mkdir -p gpurun_out
rm -f gpurun_out/tp_parity_test.log
timeout 1200 python -m pytest tests/test_gpu_tp.py -q > gpurun_out/r2_tp_tests.log 2>&1; echo "tp tests rc=$?" >> gpurun_out/r2_tp_tests.log
tail -15 gpurun_out/r2_tp_tests.log; cat gpurun_out/tp_parity_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_c2_tp2.json 2> gpurun_out/r2_bench_c2_tp2.err; echo "bench tp2 rc=$?"
head -c 3500 gpurun_out/r2_bench_c2_tp2.json; tail -5 gpurun_out/r2_bench_c2_tp2.err
SQ_TP_SHOT=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-tp-parity > gpurun_out/r2_bench_c2_tp2_shot2.json 2> gpurun_out/r2_bench_c2_tp2_shot2.err; echo "bench tp2 shot2 rc=$?"
head -c 700 gpurun_out/r2_bench_c2_tp2_shot2.json
