#!/bin/bash
run() { tag=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2v_$tag.json 2> gpurun_out/r2v_$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2v_{tag}.json").read().strip().splitlines()[-1])
    print(tag, f"{d['value']/1e9:.3f} Gpts/s", f"{d['ms_per_step']:.3f} ms", f"frac {d['roofline']['frac']:.3f}", "parity", d['parity']['ok'], d['parity'].get('worst_ratio'), d['parity'].get('error'), flush=True)
except Exception as e:
    print(tag, "FAILED", e, flush=True)
PY
}
run c5_p12_pf cfg5
run c5_p12_nopf cfg5 BSPY_EXP_B=0
run c5_p22_pf cfg5 BSPY_IMAGE=1022
run c5_p13_pf cfg5 BSPY_IMAGE=1013
run c4_p13_pf cfg4 BSPY_IMAGE=1013
run c4_p32_pf cfg4 BSPY_IMAGE=1032
run c4_p14_pf cfg4 BSPY_IMAGE=1014
