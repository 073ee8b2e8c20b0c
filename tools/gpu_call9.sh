#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/teacher_forced.log gpurun_out/logit_err.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_draft or swiglu" > gpurun_out/r2i_t.log 2>&1; rc=$?; echo "draft pytest rc=$rc"; tail -25 gpurun_out/r2i_t.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_tall.log 2>&1; echo "full pytest rc=$?"; tail -6 gpurun_out/r2i_tall.log; cat gpurun_out/logit_err.log
for v in 1 0; do
  SQ_DRAFT_FUSED=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2i_bench_d$v.json 2> gpurun_out/r2i_bench_d$v.err; echo "bench fused=$v rc=$?"
done
python - <<'PY'
import json
for f in ("d1", "d0"):
    try:
        d = json.load(open(f"gpurun_out/r2i_bench_{f}.json")); print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["accepted_tokens_per_step"], d["phases"], d["device_errors"], d["gpu_launches"])
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/r2i_bench_{f}.err").read()[-1500:])
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-cpu-baseline --no-micro --timeline gpurun_out/r2i_timeline_c2.md > gpurun_out/r2i_bench_tl.json 2> gpurun_out/r2i_bench_tl.err; echo "timeline rc=$?"; head -40 gpurun_out/r2i_timeline_c2.md | cut -c1-200
timeout 900 python bench.py --config c4 --steps 6 --warmup 3 --no-reference-gpu --no-cpu-baseline --no-micro > gpurun_out/r2i_bench_c4_1gpu.json 2> gpurun_out/r2i_bench_c4_1gpu.err; echo "c4 N=1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_c4_1gpu.json')); print(d['ms_per_step'], d['phases'])"
