#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tp_parity_test.log
timeout 1200 python -m pytest tests/test_gpu_tp.py -q > gpurun_out/r2n_tp_tests.log 2>&1; echo "tp tests rc=$?"; tail -6 gpurun_out/r2n_tp_tests.log; cat gpurun_out/tp_parity_test.log
for sh in 4 1; do
SQ_TP_SHOT=$sh timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$sh bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2n_bench_c2_tp2_shot$sh.json 2> gpurun_out/r2n_bench_c2_tp2_shot$sh.err; echo "bench tp2 shot$sh rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2n_bench_c2_tp2_shot$sh.json")); print("shot$sh", d["ms_per_step"], d["value"], d["phases"]["verify_ms_per_step"], d.get("tp_parity"), d["device_errors"], {k: v.get("us") for k, v in (d.get("kernels") or {}).items() if "tp_" in k})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2n_bench_c2_tp2_shot$sh.err").read()[-800:])
PY
done
