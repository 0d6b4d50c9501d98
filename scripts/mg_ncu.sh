#!/bin/bash
# ncu evidence for the opt-in multigrid solver (run under gpurun): launch list of one S16 solve through scripts/mg_probe.py
# (plain run first), then --set full captures of the level-0 sweep and of the outer SpMV.  Usage: scripts/mg_ncu.sh <tag>
TAG=${1:-r02}
CMD="python scripts/mg_probe.py 4000 nojac"
mkdir -p gpurun_out
$CMD > gpurun_out/mg_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/mg_plain_$TAG.log; exit 1; }
tail -3 gpurun_out/mg_plain_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/mg_launches_$TAG.csv $CMD > gpurun_out/mg_ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
# default hierarchy of S16 (10 levels, schedule (1,1,3), the three coarsest levels in k_mg_tail): 30 leg launches per cycle
# (launch 30 k is a level-0 down-leg, 30 k + 29 a level-0 up-leg) and 17 restrictions (17 k: level 0)
ncu --set full --clock-control none --import-source on -k regex:k_mg_leg -s 300 -c 1 -f -o gpurun_out/prof_mg_down_$TAG $CMD > gpurun_out/mg_ncu_full_$TAG.log 2>&1
echo "full capture (level-0 down-leg) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_mg_leg -s 329 -c 1 -f -o gpurun_out/prof_mg_up_$TAG $CMD >> gpurun_out/mg_ncu_full_$TAG.log 2>&1
echo "full capture (level-0 up-leg) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_mg_restrict -s 170 -c 1 -f -o gpurun_out/prof_mg_restrict_$TAG $CMD >> gpurun_out/mg_ncu_full_$TAG.log 2>&1
echo "full capture (level-0 restriction) rc=$?"
