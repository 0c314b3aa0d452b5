#!/bin/bash
run() { tag=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/s2v_$tag.json 2> gpurun_out/s2v_$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/s2v_{tag}.json").read().strip().splitlines()[-1])
    print(tag, f"{d['value']/1e9:.3f} Gpts/s", f"{d['ms_per_step']:.3f} ms", f"frac {d['roofline']['frac']:.3f}", "parity", d['parity']['ok'], d['parity'].get('worst_ratio'), flush=True)
except Exception as e:
    print(tag, "FAILED", e, flush=True)
    print(open(f"gpurun_out/s2v_{tag}.err").read()[-1500:])
PY
}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run c4 cfg4
run c4_noov cfg4 BSPY_BIN_OVERLAP=0
run c4soa cfg4_soa
run c5 cfg5
