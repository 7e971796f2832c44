#!/bin/bash
# ncu evidence for one round (run under gpurun on ONE B200).  usage: scripts/profile_gpu.sh <tag>
# 1) plain run (must exit 0), 2) launch list with per-launch device time, 3) one --set full capture of the two dominant kernels.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --clips 32 --steps 1 --warmup 3 --skip-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 330 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pair_linear -s 60 -c 4 -o gpurun_out/${TAG}_pair_linear $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -6
