#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step (share of each kernel), (2) --set full capture of
# the hot kernels of the headline path and of the config-5 path.  Run under gpurun on ONE GPU.  Numbers printed under
# ncu are never bench values.
set -x
OUT=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --frames-per-step 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/launches.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_svd16_l4 -s 6 -c 3 -f -o $OUT/prof_svd $CMD > $OUT/prof_svd.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval3 -s 20 -c 3 -f -o $OUT/prof_eval $CMD > $OUT/prof_eval.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_svd_warp|k_eval_c" -c 4 -f -o $OUT/prof_c5 python tools/time_config5.py 512 > $OUT/prof_c5.out 2>&1
ls -la $OUT
