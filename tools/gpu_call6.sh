#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "attention or top_p" > gpurun_out/r2f_t.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_t.log
rm -f gpurun_out/r2f_probe.log
for c in c2 c3 c4tp8 c4 c4draft; do SQ_ATTN_TIMING=1 PROBE_CFG=$c timeout 120 python tools/attn_probe.py >> gpurun_out/r2f_probe.log 2>&1; done
SQ_PDL=1 PROBE_CFG=c2 timeout 120 python tools/attn_probe.py >> gpurun_out/r2f_probe.log 2>&1
cat gpurun_out/r2f_probe.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_tall.log 2>&1; echo "full pytest rc=$?"; tail -4 gpurun_out/r2f_tall.log
