#!/bin/bash
# Round-2 profile set of the final build: per-config launch lists (the command is the bench line's own, shortened) and one
# `ncu --set full` capture of every config's dominant kernel.  Reports land in gpurun_out/; tools/ncu_summary.py and
# tools/launch_summary.py turn them into the text files under profiles/.
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 0 --no-secondary-e2e"
full() { tag=$1; kern=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -o /tmp/r02_$tag -f $B "$@" > gpurun_out/r02_${tag}_ncu.log 2>&1
  echo "$tag rc=$?"
  # the reports stay on the box (gpurun_out/ is capped at 64 MiB): the summary and the per-instruction source page travel
  python tools/ncu_summary.py /tmp/r02_$tag.ncu-rep > gpurun_out/r02_${tag}_ncu_full.txt 2>/dev/null
  ncu -i /tmp/r02_$tag.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r02_${tag}_source.csv.gz; }
list() { tag=$1; shift
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_${tag}_launches.csv $B "$@" > /dev/null 2>&1
  echo "$tag launches rc=$?"; }
list cfg2 --config cfg2
list cfg1 --config cfg1
list cfg3 --config cfg3 --scale 0.2
list cfg4 --config cfg4 --scale 0.34
list cfg5 --config cfg5 --scale 0.27
list grid3 --config grid3
full cfg2_grid2_dmma grid2_dmma 3 --config cfg2 --scale 0.5
full cfg1_curve_tab_poly eval_curve_tab 3 --config cfg1 --scale 20
full cfg3_many_tab many_tab 2 --config cfg3 --scale 0.05
full cfg4_eval_staged2 eval_staged2 6 --config cfg4 --scale 0.17
full cfg5_eval_image2 eval_image2 6 --config cfg5 --scale 0.14
full grid3_dmma grid3_dmma 3 --config grid3 --scale 0.5
