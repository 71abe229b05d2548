#!/bin/bash
# round-2 final single-GPU evidence (under gpurun, 1 B200): smoke, pytest -m gpu, bench, reference arm, ncu step capture + launch list, sweeps
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc $?"
python benchmarks/profile_step.py > gpurun_out/r02_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_step_final python benchmarks/profile_step.py > gpurun_out/r02_ncu.log 2>&1
echo "ncu rc $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train --repeats 0 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train --repeats 0 > gpurun_out/r02_launches.log 2>&1
echo "launch list rc $?"
python benchmarks/sort_bench.py > gpurun_out/r02_sort_bench.jsonl 2> gpurun_out/r02_sort_bench.err; echo "sort rc $?"
python benchmarks/render_sweep.py > gpurun_out/r02_render_sweep.txt 2> gpurun_out/r02_render_sweep.err; echo "sweep rc $?"
python benchmarks/data_side_bench.py > gpurun_out/r02_data_side.jsonl 2> gpurun_out/r02_data_side.err; echo "data rc $?"
python - <<PY
import json
b=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print(b["value"], b["ms_per_step"], b["repeats"]["ms_per_step_median"], "e2e", b["e2e"]["ms_per_step"], b["e2e_fused_loss"]["ms_per_step"], "train", b["train"]["ms_per_step"], b["gpu_launches_per_step"])
print({k:round(x,4) for k,x in b["stage_ms"].items()})
for s in b["stages"]: print(s["kernel"], round(s["frac"],3), s.get("ms"))
PY
