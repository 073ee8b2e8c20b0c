#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/teacher_forced.log gpurun_out/logit_err.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "attention" -x -q > gpurun_out/r2b_attn_tests.log 2>&1; rc=$?; echo "attn tests rc=$rc" >> gpurun_out/r2b_attn_tests.log
tail -5 gpurun_out/r2b_attn_tests.log
for c in c2 c3 c4tp8 c4 c4draft; do SQ_ATTN_TIMING=1 PROBE_CFG=$c timeout 120 python tools/attn_probe.py >> gpurun_out/r2b_probe.log 2>&1; done
cat gpurun_out/r2b_probe.log
if [ $rc -eq 0 ]; then
  timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_t.log
  tail -15 gpurun_out/r2b_t.log
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench rc=$?" >> gpurun_out/r2b_bench_c2.err
  head -c 2500 gpurun_out/r2b_bench_c2.json; tail -3 gpurun_out/r2b_bench_c2.err
  SQ_PDL=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -k "attention" -x -q > gpurun_out/r2b_attn_tests_pdl.log 2>&1; echo "pdl attn rc=$?"
  SQ_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2b_bench_c2_pdl.json 2> gpurun_out/r2b_bench_c2_pdl.err; echo "pdl bench rc=$?"
  head -c 600 gpurun_out/r2b_bench_c2_pdl.json; tail -3 gpurun_out/r2b_bench_c2_pdl.err
fi
