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
# GEMM v2 sweep: default tile choice + forced alternatives (bn,split,mc)
timeout 300 python tools/gemm_probe.py > gpurun_out/r2b_gemm_default.log 2>&1
for f in "qkv:128,1,2" "qkv:192,1,2" "qkv:96,1,1" "gate_up:256,1,2" "gate_up:160,1,1" "gate_up:128,1,1" "o:128,4,1" "o:64,4,2" "o:64,2,2" "down:128,4,1" "down:64,4,2" "lm_head:256,1,1"; do
  SQ_GEMM_FORCE=${f#*:} PROBE_ONLY=${f%%:*} timeout 120 python tools/gemm_probe.py >> gpurun_out/r2b_gemm_forced.log 2>&1
done
SQ_PDL=1 timeout 300 python tools/gemm_probe.py > gpurun_out/r2b_gemm_pdl.log 2>&1
cat gpurun_out/r2b_gemm_default.log gpurun_out/r2b_gemm_forced.log gpurun_out/r2b_gemm_pdl.log | cut -c1-230
# the reference's own tests/testbed.py, verbatim, on the drop-in modules (c2 shapes, bundled openwebtext prompts)
timeout 600 python tools/run_reference_testbed.py -- --model random-init:llama-68m:1 --target random-init:llama-2-7b:2 --growmap $PWD/A100_growmaps/68m_7b/growmaps/A100-CNN-68m-7b-stochastic.pt --T 0.6 --P 1.0 --M 384 --dataset openwebtext --start 0 --end 20 --Mode greedy > gpurun_out/r2b_ref_testbed_verbatim.log 2>&1; echo "verbatim testbed rc=$?"
tail -4 gpurun_out/r2b_ref_testbed_verbatim.log
