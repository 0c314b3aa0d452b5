#!/bin/bash
# usage: tools/run_cfg1.sh <tag> VAR=val ...  -> cfg1 at 1e6 and 1e8 points
tag=$1; shift
for sc in 1 100; do
env "$@" python bench.py --config cfg1 --scale $sc --steps 20 --warmup 5 --cpu-seconds 0.3 --e2e-steps 1 > gpurun_out/${tag}_x$sc.json 2> gpurun_out/${tag}_x$sc.err
python - "${tag}_x$sc" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
    print(tag, f"{d['value']/1e9:.2f} Gpts/s", f"{d['ms_per_step']*1e3:.1f} us")
except Exception as e:
    print(tag, "FAILED", e)
PY
done
