#!/bin/bash
# final 1-GPU validation of HEAD: full GPU suite, smoke(), default bench line, ncu capture of the attention kernel + launch list
mkdir -p gpurun_out
rm -f gpurun_out/teacher_forced.log gpurun_out/logit_err.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2z_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_t.log; tail -4 gpurun_out/r2z_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2z_bench_c2.json 2> gpurun_out/r2z_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2z_bench_c2.json"))
print(d["ms_per_step"], d["value"], d["e2e"], d["roofline"]["frac"], d["roofline"]["traffic"], d["phases"], (d.get("reference_gpu") or {}).get("value"), (d.get("reference_gpu") or {}).get("speedup_vs_reference_gpu"), d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"])
PY
PROBE_L=4 timeout 200 ncu --set full --clock-control none --import-source on -k regex:tree_attn_tc -s 10 -c 1 -f -o gpurun_out/r2z_attn_prof python tools/attn_probe.py > gpurun_out/r2z_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-micro --no-reference-gpu --no-cpu-baseline > gpurun_out/r2z_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
