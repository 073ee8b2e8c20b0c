#!/bin/bash
# 1-GPU call: full GPU suite, PDL variants, ncu (GEMM ours vs cuBLASLt, attention, launch list), bench c3 / c5, verbatim testbed
mkdir -p gpurun_out
rm -f gpurun_out/teacher_forced.log gpurun_out/logit_err.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_t.log
tail -12 gpurun_out/r2c_t.log; cat gpurun_out/logit_err.log
SQ_PDL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode.py -q -x > gpurun_out/r2c_t_pdl.log 2>&1; echo "pdl pytest rc=$?"; tail -3 gpurun_out/r2c_t_pdl.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err; echo "bench rc=$?"
SQ_PDL=1 timeout 600 python bench.py --steps 40 --warmup 5 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2c_bench_c2_pdl.json 2> gpurun_out/r2c_bench_c2_pdl.err; echo "pdl bench rc=$?"
python - <<'PY'
import json
for f in ("r2c_bench_c2", "r2c_bench_c2_pdl"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["value"], d["roofline"]["us_per_launch"], d["kernels"])
    except Exception as e: print(f, "ERR", e)
PY
# ncu: our GEMM vs the cuBLASLt kernel on the qkv shape (full sections), attention full, launch list of a c2 step
PROBE_ONLY=qkv SQ_GEMM_FORCE=96,1,2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_kernel|nvjet|gemm" -c 4 -f -o gpurun_out/r2c_gemm_qkv python tools/gemm_probe.py > gpurun_out/r2c_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
PROBE_ONLY=o SQ_GEMM_FORCE=128,4,1 timeout 300 ncu --set full --clock-control none -k regex:"gemm_tn_kernel|nvjet|gemm" -c 4 -f -o gpurun_out/r2c_gemm_o python tools/gemm_probe.py > gpurun_out/r2c_ncu_gemm_o.log 2>&1; echo "ncu gemm o rc=$?"
PROBE_L=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tree_attn_tc -s 10 -c 1 -f -o gpurun_out/r2c_attn_prof python tools/attn_probe.py > gpurun_out/r2c_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-micro --no-reference-gpu --no-cpu-baseline > gpurun_out/r2c_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py --config c3 --steps 40 --warmup 5 > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err; echo "c3 rc=$?"; head -c 400 gpurun_out/r2c_bench_c3.json
timeout 900 python bench.py --config c5 --steps 12 --warmup 3 > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err; echo "c5 rc=$?"; head -c 300 gpurun_out/r2c_bench_c5.json
timeout 600 python tools/run_reference_testbed.py -- --model random-init:llama-68m:1 --target random-init:llama-2-7b:2 --growmap $PWD/A100_growmaps/68m_7b/growmaps/A100-CNN-68m-7b-stochastic.pt --T 0.6 --P 1.0 --M 384 --dataset openwebtext --start 0 --end 20 --Mode greedy > gpurun_out/r2c_ref_testbed_verbatim.log 2>&1; echo "verbatim testbed rc=$?"
tail -6 gpurun_out/r2c_ref_testbed_verbatim.log | cut -c1-300
