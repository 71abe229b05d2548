#!/bin/bash
# usage (under gpurun --gpus N): bash scripts/run_scaling.sh N   -- bench + configs[2] / configs[3] trainer runs at N GPUs
N=$1
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
O=gpurun_out
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 $2 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err; echo "bench rc $?"
timeout 600 $TR benchmarks/train_step.py --scene s2 --gaussians 3000000 --steps 30 > $O/r02_train_s2_n$N.json 2> $O/r02_train_s2_n$N.err; echo "s2 rc $?"
timeout 900 $TR benchmarks/train_step.py --scene s3 --gaussians 6000000 --refine-every 100 --start-step 600 --steps 210 > $O/r02_train_s3_n$N.json 2> $O/r02_train_s3_n$N.err; echo "s3 rc $?"
if [ "$N" != "1" ]; then
timeout 600 $TR benchmarks/train_step.py --scene s2 --gaussians 3000000 --steps 30 --comm nccl > $O/r02_train_s2_n${N}_nccl.json 2> $O/r02_train_s2_n${N}_nccl.err; echo "s2 nccl rc $?"
fi
python - <<PY
import json
for f in ("bench_n$N","train_s2_n$N","train_s3_n$N","train_s2_n${N}_nccl"):
    try:
        b=json.loads(open("$O/r02_%s.json"%f).read().strip().splitlines()[-1])
        if "value" in b: print(f, b["value"], b["ms_per_step"], "e2e", b["e2e"]["ms_per_step"], b["e2e_fused_loss"]["ms_per_step"], "train", b["train"]["ms_per_step"], (b.get("multi_gpu_check") or {}).get("ok_all_ranks"))
        else: print(f, {k:b[k] for k in ("ms_per_step","train_iters_per_s","gaussians_end","comm","replicas_bit_identical","n_refines_in_window","overflow_repeats") if k in b})
    except Exception as e: print(f, "ERR", e)
PY
