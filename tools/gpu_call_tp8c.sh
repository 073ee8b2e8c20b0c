#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; np=$2; port=$3; shift 3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 40 --warmup 5 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$name.json")); print("$name", d["ms_per_step"], d["value"], d["phases"]["verify_ms_per_step"], (d.get("tp_parity") or {}).get("accept_seq_identical_steps"), d["device_errors"], {k: v.get("us") for k, v in (d.get("kernels") or {}).items() if "tp_" in k})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/$name.err").read()[-800:])
PY
}
run r2p_c2_tp8_ll 8 29561
SQ_TP_SHOT=2 run r2p_c2_tp8_pull2 8 29562 --no-tp-parity --no-micro
run r2p_c2_tp4_ll 4 29563
SQ_TP_SHOT=2 run r2p_c2_tp4_pull2 4 29564 --no-tp-parity --no-micro
