#!/bin/bash
# Evidence of the final build on ONE B200 (run under gpurun): the default bench line, the ncu launch lists of the inference forward and of one
# training step, every kernel row alone.  usage: scripts/final_evidence.sh <tag>
set -u
TAG=${1:-r2f}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc=$?"
CMD="python bench.py --clips 128 --steps 1 --warmup 3 --headline-only --skip-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 345 -c 230 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_summary.md
python scripts/train_launches.py > /dev/null 2>&1 && \
ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_launches.csv python scripts/train_launches.py > gpurun_out/${TAG}_train_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches.csv > gpurun_out/${TAG}_train_launch_summary.md
python scripts/kernel_bench.py 128 > gpurun_out/${TAG}_kernel_rooflines.json 2> gpurun_out/${TAG}_kb.err
python scripts/membound_bench.py 2> gpurun_out/${TAG}_mb.err | grep -v "^{" > gpurun_out/${TAG}_membound_rows.txt
python scripts/segments_bench.py 128 > gpurun_out/${TAG}_segments_rows.json 2> gpurun_out/${TAG}_sb.err
head -12 gpurun_out/${TAG}_launch_summary.md
