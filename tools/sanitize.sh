#!/bin/bash
# compute-sanitizer passes over tools/sanitize_workload.py; logs land in gpurun_out/sanitize_<tool>.log
# usage (GPU box): bash tools/sanitize.sh [n_photon]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-20000}
for tool in memcheck racecheck synccheck initcheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_workload.py $N > gpurun_out/sanitize_$tool.log 2>&1
    echo "== $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
