#!/bin/bash
# A B A B of the default inference bench line under an environment switch: scripts/ab_env.sh VAR valueA valueB [tag]
var=$1; a=$2; b=$3; tag=${4:-ab}
for i in 1 2; do
for v in $a $b; do
env $var=$v python bench.py --headline-only --skip-cpu-baseline --steps 4 --warmup 3 > gpurun_out/${tag}_${v}_$i.json 2> gpurun_out/${tag}_${v}_$i.err
done; done
python - "$tag" <<'P'
import json,glob,sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        fam=d["roofline"]["families"]
        print(f, round(d["value"]), d["clocks"]["sm_mhz"], {k:(round(v["avg_us"],1), round(v["share_of_step"],3)) for k,v in fam.items()})
    except Exception as e: print(f, "ERR", e)
P
