#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/logit_err.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "fused_draft" > gpurun_out/r2k_t.log 2>&1; rc=$?; echo "draft pytest rc=$rc"; tail -12 gpurun_out/r2k_t.log | cut -c1-250; cat gpurun_out/logit_err.log
for v in chain coop 0; do
  SQ_DRAFT_FUSED=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline --no-micro > gpurun_out/r2k_bench_d$v.json 2> gpurun_out/r2k_bench_d$v.err; echo "bench fused=$v rc=$?"
done
python - <<'PY'
import json
for f in ("dchain", "dcoop", "d0"):
    try:
        d = json.load(open(f"gpurun_out/r2k_bench_{f}.json")); print(f, d["ms_per_step"], d["value"], d["config"]["accepted_tokens_per_step"], d["phases"]["draft_ms_per_step"], d["phases"]["verify_ms_per_step"], d["device_errors"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/r2k_bench_{f}.err").read()[-1500:])
PY
SQ_DRAFT_FUSED=chain timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-cpu-baseline --no-micro --timeline gpurun_out/r2k_timeline_chain.md > /dev/null 2>&1; head -12 gpurun_out/r2k_timeline_chain.md | cut -c1-160
SQ_DRAFT_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-cpu-baseline --no-micro --timeline gpurun_out/r2k_timeline_multi.md > /dev/null 2>&1; head -40 gpurun_out/r2k_timeline_multi.md | cut -c1-160
