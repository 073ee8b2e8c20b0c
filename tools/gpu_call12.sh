#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/logit_err.log gpurun_out/teacher_forced.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2l_tall.log 2>&1; echo "full pytest rc=$?"; tail -6 gpurun_out/r2l_tall.log; cat gpurun_out/logit_err.log gpurun_out/teacher_forced.log
for v in 1 0; do
  SQ_DRAFT_ATTN=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline --no-micro > gpurun_out/r2l_bench_a$v.json 2> gpurun_out/r2l_bench_a$v.err; echo "bench draft_attn=$v rc=$?"
done
python - <<'PY'
import json
for f in ("a1", "a0"):
    try:
        d = json.load(open(f"gpurun_out/r2l_bench_{f}.json")); print(f, d["ms_per_step"], d["value"], d["config"]["accepted_tokens_per_step"], d["phases"]["draft_ms_per_step"], d["phases"]["verify_ms_per_step"], d["device_errors"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/r2l_bench_{f}.err").read()[-1500:])
PY
timeout 600 python bench.py --config c3 --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline --no-micro > gpurun_out/r2l_bench_c3.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_c3.json')); print('c3', d['ms_per_step'], d['value'], d['phases'])"
