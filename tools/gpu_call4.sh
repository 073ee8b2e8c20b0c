#!/bin/bash
# 1-GPU call: tests, GEMM v3 (decoupled rings) sweep, bench with SQ_GEMM=1 (+PDL)
mkdir -p gpurun_out
rm -f gpurun_out/teacher_forced.log gpurun_out/logit_err.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2d_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_t.log
tail -8 gpurun_out/r2d_t.log; cat gpurun_out/logit_err.log
timeout 300 python tools/gemm_probe.py > gpurun_out/r2d_gemm_default.log 2>&1
for f in "qkv:96,1,2" "qkv:96,1,1" "qkv:128,1,2" "qkv:128,1,1" "qkv:192,1,2" "gate_up:160,1,2" "gate_up:160,1,1" "gate_up:192,1,1" "gate_up:224,1,1" "gate_up:256,1,2" "o:128,4,1" "o:128,2,2" "o:64,2,1" "o:128,2,1" "o:64,2,2" "down:128,4,1" "down:128,4,2" "down:128,2,1" "down:64,2,1" "lm_head:224,1,1" "lm_head:256,1,1" "lm_head:256,1,2"; do
  SQ_GEMM_FORCE=${f#*:} PROBE_ONLY=${f%%:*} timeout 120 python tools/gemm_probe.py >> gpurun_out/r2d_gemm_forced.log 2>&1
done
SQ_PDL=1 timeout 300 python tools/gemm_probe.py > gpurun_out/r2d_gemm_pdl.log 2>&1
cat gpurun_out/r2d_gemm_default.log gpurun_out/r2d_gemm_forced.log gpurun_out/r2d_gemm_pdl.log | cut -c1-230
SQ_GEMM=1 timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_kernels.py -q -x > gpurun_out/r2d_t_gemm.log 2>&1; echo "gemm pytest rc=$?"; tail -3 gpurun_out/r2d_t_gemm.log
for v in "0 0" "1 0" "1 1" "0 1"; do set -- $v
  SQ_GEMM=$1 SQ_PDL=$2 timeout 600 python bench.py --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2d_bench_g$1p$2.json 2> gpurun_out/r2d_bench_g$1p$2.err; echo "bench gemm=$1 pdl=$2 rc=$?"
done
python - <<'PY'
import json
for f in ("g0p0", "g1p0", "g1p1", "g0p1"):
    try:
        d = json.load(open(f"gpurun_out/r2d_bench_{f}.json")); print(f, d["ms_per_step"], d["value"], d["roofline"]["us_per_launch"], d["device_errors"])
    except Exception as e: print(f, "ERR", e)
PY
