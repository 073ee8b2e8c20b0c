#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/teacher_forced.log gpurun_out/logit_err.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2g_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_t.log
tail -8 gpurun_out/r2g_t.log; cat gpurun_out/logit_err.log gpurun_out/teacher_forced.log
for v in "auto 0" "auto 1" "0 0"; do set -- $v
  SQ_GEMM=$1 SQ_PDL=$2 timeout 600 python bench.py --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2g_bench_g$1p$2.json 2> gpurun_out/r2g_bench_g$1p$2.err; echo "bench gemm=$1 pdl=$2 rc=$?"
done
python - <<'PY'
import json
for f in ("gautop0", "gautop1", "g0p0"):
    try:
        d = json.load(open(f"gpurun_out/r2g_bench_{f}.json")); print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["us_per_launch"], d["device_errors"], d["gpu_launches"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/r2g_bench_{f}.err").read()[-1500:])
PY
SQ_PDL=1 timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_kernels.py -q -x > gpurun_out/r2g_t_pdl.log 2>&1; echo "pdl pytest rc=$?"; tail -3 gpurun_out/r2g_t_pdl.log
