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
timeout 900 python -m pytest tests -m gpu -x -q -k "cell_kernel" 2>&1 | tail -5
run c5_s2_22 cfg5 BSPY_CELL_POLY=2022
run c5_s2_32 cfg5 BSPY_CELL_POLY=2032
run c5_s2_22_ov cfg5 BSPY_CELL_POLY=2022 BSPY_BIN_OVERLAP=1
run c5_s2_62 cfg5 BSPY_CELL_POLY=2062
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --no-secondary-e2e"
BSPY_CELL_POLY=2022 ncu --set full --clock-control none --import-source on -k regex:eval_poly_s2 -s 6 -c 1 -o /tmp/r02_cfg5_s2 -f $B --config cfg5 --scale 0.14 > gpurun_out/r02_cfg5_s2_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02_cfg5_s2.ncu-rep > gpurun_out/r02_cfg5_eval_poly_s2_ncu_full.txt
ncu -i /tmp/r02_cfg5_s2.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r02_cfg5_eval_poly_s2_source.csv.gz
