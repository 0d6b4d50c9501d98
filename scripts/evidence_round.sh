#!/bin/bash
# One GPU box, one call: GPU test suite, default bench (both arms), the extra single-GPU records and the ncu round.
# Usage (under gpurun): bash scripts/evidence_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 60 python scripts/extra_records.py 300 > gpurun_out/extra_small.log 2>&1; echo "extra-small rc=$?"; tail -2 gpurun_out/extra_small.log | cut -c1-300
timeout 200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 100 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2>/dev/null; echo "reference arm rc=$?"
timeout 150 python scripts/extra_records.py > gpurun_out/extra_records_$TAG.jsonl 2> gpurun_out/extra_records_$TAG.err; echo "extra rc=$?"
cut -c1-420 gpurun_out/extra_records_$TAG.jsonl
timeout 420 bash scripts/ncu_round.sh $TAG
