#!/bin/bash
run() { tag=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2v_$tag.json 2> gpurun_out/r2v_$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2v_{tag}.json").read().strip().splitlines()[-1])
    print(tag, f"{d['value']/1e9:.3f} Gpts/s", f"{d['ms_per_step']:.3f} ms", f"frac {d['roofline']['frac']:.3f}", "parity", d['parity']['ok'], d['parity'].get('worst_ratio'), flush=True)
except Exception as e:
    print(tag, "FAILED", e, flush=True)
PY
}
python -m pytest tests -m gpu -x -q -k "record_mode or cell_kernel or binned" 2>&1 | tail -3
run c5_img cfg5
run c5_img14 cfg5 BSPY_IMAGE=14
run c5_img123 cfg5 BSPY_IMAGE=123
run c5_img132 cfg5 BSPY_IMAGE=132
run c5_img113 cfg5 BSPY_IMAGE=113
run c5_img115 cfg5 BSPY_IMAGE=115
run c5_img62 cfg5 BSPY_IMAGE=62
run c5_img_ov cfg5 BSPY_BIN_OVERLAP=1
run c5_img_c23 cfg5 BSPY_BIN_REC_CHUNK_LOG2=23
run c5soa_img cfg5_soa
run c4_img34 cfg4 BSPY_IMAGE=34
run c4_img33 cfg4 BSPY_IMAGE=33
run c4_img34_ov cfg4 BSPY_IMAGE=34 BSPY_BIN_OVERLAP=1
