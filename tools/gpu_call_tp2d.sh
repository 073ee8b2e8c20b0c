#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tp_parity_test.log
timeout 1200 python -m pytest tests/test_gpu_tp.py -q > gpurun_out/r2o_tp_tests.log 2>&1; echo "tp tests rc=$?"; tail -4 gpurun_out/r2o_tp_tests.log; cat gpurun_out/tp_parity_test.log
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 40 --warmup 5 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$name.json")); print("$name", d["ms_per_step"], d["value"], d["phases"]["verify_ms_per_step"], (d.get("tp_parity") or {}).get("accept_seq_identical_steps"), d["device_errors"], {k: v.get("us") for k, v in (d.get("kernels") or {}).items() if "tp_" in k})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$name.err").read()[-800:])
PY
}
PORT=29551 run r2o_tp2_ll_msgll
PORT=29552 SQ_TP_MSG=nccl run r2o_tp2_ll_msgnccl --no-tp-parity
PORT=29553 SQ_TP_SHOT=1 SQ_TP_MSG=nccl run r2o_tp2_pull_msgnccl --no-tp-parity
