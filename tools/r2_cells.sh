#!/bin/bash
# round-2 experiment driver for the cell-sorted path (run under gpurun): launch lists + one full capture per shape
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-graph --e2e-steps 1"
for cfg in cfg4 cfg5; do
  for ck in 1 0; do
    BSPY_CELL_KERNEL=$ck python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r2_${cfg}_ck$ck.json 2> gpurun_out/r2_${cfg}_ck$ck.err
  done
  BSPY_CELL_KERNEL=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r2_${cfg}_ck1_launches.csv $B --config $cfg --scale 0.34 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:eval_cell_mma -s 3 -c 1 -o gpurun_out/r2_${cfg}_cell -f $B --config $cfg --scale 0.2 > gpurun_out/r2_${cfg}_ncu.log 2>&1
done
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
wl = bench.Cfg2Teapot(); wl.setup(torch.device('cuda:0'), 0, 1.0)
for i in range(4):
    torch.cuda.synchronize(); t = time.perf_counter(); wl.e2e_step(); torch.cuda.synchronize(); print('cfg2 e2e step', i, time.perf_counter() - t, flush=True)
PY
