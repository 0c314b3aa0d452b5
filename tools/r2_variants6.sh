#!/bin/bash
run() { tag=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2v_$tag.json 2> gpurun_out/r2v_$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2v_{tag}.json").read().strip().splitlines()[-1])
    print(tag, f"{d['value']/1e9:.3f} Gpts/s", f"{d['ms_per_step']*1e3:.2f} us", f"frac {d['roofline']['frac']:.3f}", "parity", d['parity']['ok'], d['parity'].get('worst_ratio'), d['parity'].get('error'), flush=True)
except Exception as e:
    print(tag, "FAILED", e, flush=True)
PY
}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run c3_poly cfg3
run c3_rec cfg3 BSPY_MANY_MODE=0
run c1_u8 cfg1
run c1_u2 cfg1 BSPY_EXP_A=1
run g3_default grid3
run g3_chunk1024 grid3 BSPY_GRID3_CHUNK=1024
