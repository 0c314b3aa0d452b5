#!/bin/bash
# usage: tools/r2_variants.sh  -> one bench line per variant into gpurun_out/r2v_<tag>.json, summary on stdout
run() { tag=$1; cfg=$2; shift 2; env "$@" python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r2v_$tag.json 2> gpurun_out/r2v_$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2v_{tag}.json").read().strip().splitlines()[-1])
    print(tag, f"{d['value']/1e9:.3f} Gpts/s", f"{d['ms_per_step']:.3f} ms", f"frac {d['roofline']['frac']:.3f}", flush=True)
except Exception as e:
    print(tag, "FAILED", e, flush=True)
PY
}
python -m pytest tests -m gpu -x -q -k "record_mode or cell_kernel or binned" 2>&1 | tail -3
run c4_default cfg4
run c4_v1 cfg4 BSPY_EXP_A=1
run c4_s33 cfg4 BSPY_STAGED=33
run c4_s33_noov cfg4 BSPY_STAGED=33 BSPY_BIN_OVERLAP=0
run c4_noov cfg4 BSPY_BIN_OVERLAP=0
run c4_chunk23 cfg4 BSPY_BIN_REC_CHUNK_LOG2=23
run c4_chunk24 cfg4 BSPY_BIN_REC_CHUNK_LOG2=24
run c4_chunk21 cfg4 BSPY_BIN_REC_CHUNK_LOG2=21
run c4soa_default cfg4_soa
run c5_default cfg5
run c5_t0 cfg5 BSPY_DEP_TILE=0
run c5_t24 cfg5 BSPY_DEP_TILE=24
run c5_ov1 cfg5 BSPY_BIN_OVERLAP=1
run c5_s14 cfg5 BSPY_STAGED=14
run c5_s23 cfg5 BSPY_STAGED=23
run c5_s62 cfg5 BSPY_STAGED=62
run c5_chunk24 cfg5 BSPY_BIN_REC_CHUNK_LOG2=24
run c5_chunk24_s14 cfg5 BSPY_BIN_REC_CHUNK_LOG2=24 BSPY_STAGED=14
run c5soa_default cfg5_soa
