#!/bin/bash
# Multi-GPU evidence of one round (run under `gpurun --gpus N`): peer-memory kernels vs the NCCL path and the closed form,
# the headline bench through FEMSolver(partition='rows'), and config 5 (transient, theta 0 and 1/2).
# Usage: bash scripts/multi_gpu_round.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
timeout 240 $RUN scripts/peer_check.py > gpurun_out/${TAG}_peer_check_${N}gpu.json 2> gpurun_out/${TAG}_peer_check_${N}gpu.err; echo "check rc=$?"; tail -1 gpurun_out/${TAG}_peer_check_${N}gpu.json | cut -c1-400
timeout 300 $RUN bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print(d["value"], d.get("solver"), d.get("phases_ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("verify") or {}).get("true_relative_residual"))
except Exception as e:
    print("no bench line:", e)
PY
for TH in 0 0.5; do
  timeout 300 $RUN bench.py --gpus $N --workload transient --theta $TH --steps 2 --warmup 1 > gpurun_out/${TAG}_transient_${N}gpu_theta$TH.json 2> gpurun_out/${TAG}_transient_${N}gpu_theta$TH.err; echo "transient theta=$TH rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_transient_${N}gpu_theta$TH.json").read().strip().splitlines()[-1])
    print(d["value"], d.get("solver") or d.get("transient"))
except Exception as e:
    print("no transient line:", e)
PY
done
