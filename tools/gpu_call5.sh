#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/r2e_t.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_t.log
rm -f gpurun_out/r2e_gemm.log
for t in 0 1; do
  PROBE_TILED=$t timeout 300 python tools/gemm_probe.py >> gpurun_out/r2e_gemm.log 2>&1
done
for f in "qkv:96,1,2" "qkv:96,1,1" "qkv:128,1,1" "qkv:192,1,2" "gate_up:160,1,1" "gate_up:256,1,2" "gate_up:192,1,1" "o:128,4,1" "o:64,2,1" "down:128,4,1" "lm_head:256,1,1"; do
  PROBE_TILED=1 SQ_GEMM_FORCE=${f#*:} PROBE_ONLY=${f%%:*} timeout 120 python tools/gemm_probe.py >> gpurun_out/r2e_gemm.log 2>&1
done
cut -c1-220 gpurun_out/r2e_gemm.log
