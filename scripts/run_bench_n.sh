#!/bin/bash
# usage (under gpurun --gpus N): bash scripts/run_bench_n.sh N  -- the driver's bench command at N GPUs -> gpurun_out/r02_bench_nN.json
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc $?"
python - <<PY
import json
b=json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
print($N, b["value"], b["ms_per_step"], b["repeats"]["ms_per_step_median"], "e2e", b["e2e"]["ms_per_step"], b["e2e_fused_loss"]["ms_per_step"], "train", b["train"]["ms_per_step"], b["multi_gpu_check"]["ok_all_ranks"], b["comm"])
print({k:round(x,4) for k,x in b["stage_ms"].items()})
PY
