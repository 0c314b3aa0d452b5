#!/bin/bash
# usage: tools/run_variants.sh <cfg> <scale> <tag> VAR=val ...   -> one bench line per call into gpurun_out/<tag>.json
cfg=$1; scale=$2; tag=$3; shift 3
env "$@" python bench.py --config $cfg --scale $scale --steps 3 --warmup 3 --cpu-seconds 0.5 --e2e-steps 1 > gpurun_out/$tag.json 2> gpurun_out/$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
    print(tag, f"{d['value']/1e9:.3f} Gpts/s", f"{d['ms_per_step']:.3f} ms")
except Exception as e:
    print(tag, "FAILED", e)
PY
