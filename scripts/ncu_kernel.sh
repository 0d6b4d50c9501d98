#!/bin/bash
# one --set full capture of a kernel from an arbitrary short command.  Usage: scripts/ncu_kernel.sh <tag> <kernel-regex> <skip> <cmd...>
TAG=$1; KRE=$2; SKIP=$3; shift 3
"$@" > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 2 -f -o gpurun_out/prof_$TAG "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
