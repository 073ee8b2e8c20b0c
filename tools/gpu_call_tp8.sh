#!/bin/bash
# 8-GPU call: config c4 (7B -> 70B, 768-node tree, M=1024, target TP-8) and config c2 at TP-8 / TP-4, each with tp_parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2_tp8_gpus.txt 2>&1
run() { # name, nproc, port, extra args...
  name=$1; np=$2; port=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  head -c 1800 gpurun_out/$name.json; echo; tail -3 gpurun_out/$name.err | cut -c1-300
}
run r2_bench_c4_tp8 8 29521 --config c4 --steps 10 --warmup 3
run r2_bench_c2_tp8 8 29522 --steps 40 --warmup 5
run r2_bench_c2_tp4 4 29523 --steps 40 --warmup 5
SQ_TP_SHOT=1 run r2_bench_c2_tp8_shot1 8 29524 --steps 40 --warmup 5 --no-tp-parity
