#!/bin/bash
# ncu evidence for one round (run under gpurun): launch list of a short bench run + --set full captures of the dominant
# kernel (SpMV inside the solver) and of the numeric assembly.  Usage: scripts/ncu_round.sh <tag> [spmv-regex] [asm-regex]
TAG=${1:-r01}; KRE=${2:-k_spmv_pat}; ARE=${3:-k_assemble_grid}
CMD="python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --maxiter 60 --check-every 20"
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$TAG.log; exit 1; }
# launch list: exactly the launches of the timed step (cudaProfilerStart/Stop around it), with enough iterations for the
# kernels' SHARES of the step to be those of a full solve (the symbolic phase and the warm-up step are not listed)
LCMD="python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --maxiter ${LIST_ITERS:-150} --check-every 50 --profiler-range"
$LCMD > gpurun_out/ncu_plain_list_$TAG.log 2>&1 || { echo "plain run of the launch-list command failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv $LCMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 40 -c 3 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture (spmv) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$ARE -s 1 -c 2 -f -o gpurun_out/prof_${TAG}_asm $CMD > gpurun_out/ncu_full_${TAG}_asm.log 2>&1
echo "full capture (assembly) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_vec4 -s 40 -c 2 -f -o gpurun_out/prof_${TAG}_vec $CMD > gpurun_out/ncu_full_${TAG}_vec.log 2>&1
echo "full capture (vector kernels) rc=$?"
ls -la gpurun_out | tail -10
