#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/logit_err.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "fused_draft" > gpurun_out/r2j_t.log 2>&1; rc=$?; echo "draft pytest rc=$rc"; tail -12 gpurun_out/r2j_t.log | cut -c1-250; cat gpurun_out/logit_err.log
for v in 1 0; do
  SQ_DRAFT_FUSED=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline --no-micro > gpurun_out/r2j_bench_d$v.json 2> gpurun_out/r2j_bench_d$v.err; echo "bench fused=$v rc=$?"
done
python - <<'PY'
import json
for f in ("d1", "d0"):
    try:
        d = json.load(open(f"gpurun_out/r2j_bench_{f}.json")); print(f, d["ms_per_step"], d["value"], d["config"]["accepted_tokens_per_step"], d["phases"]["draft_ms_per_step"], d["phases"]["verify_ms_per_step"], d["device_errors"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/r2j_bench_{f}.err").read()[-1500:])
PY
