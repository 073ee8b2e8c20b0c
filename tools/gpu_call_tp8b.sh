#!/bin/bash
# 8-GPU call: scaling curve of c2 at N = 2, 4, 8 (phases + tp_parity), c4 at TP-8 with the phase split, in-graph timelines on rank 0
mkdir -p gpurun_out
run() { # name, nproc, port, extra args...
  name=$1; np=$2; port=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$name.json")); print("$name", d["ms_per_step"], d["value"], d.get("phases", {}).get("draft_ms_per_step"), d.get("phases", {}).get("verify_ms_per_step"), d.get("tp_parity"), d["device_errors"], {k: v.get("us") for k, v in (d.get("kernels") or {}).items() if "tp_" in k})
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/$name.err").read()[-800:])
PY
}
run r2m_bench_c2_tp8 8 29531 --steps 40 --warmup 5
run r2m_bench_c2_tp4 4 29532 --steps 40 --warmup 5
run r2m_bench_c2_tp2 2 29533 --steps 40 --warmup 5
run r2m_bench_c4_tp8 8 29534 --config c4 --steps 10 --warmup 3
run r2m_tl_c2_tp8 8 29535 --steps 20 --warmup 5 --no-tp-parity --no-micro --timeline gpurun_out/r2m_timeline_c2_tp8.md
run r2m_tl_c4_tp8 8 29536 --config c4 --steps 6 --warmup 3 --no-tp-parity --no-micro --timeline gpurun_out/r2m_timeline_c4_tp8.md
head -25 gpurun_out/r2m_timeline_c2_tp8.md | cut -c1-150; head -25 gpurun_out/r2m_timeline_c4_tp8.md | cut -c1-150
