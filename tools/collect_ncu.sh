#!/bin/bash
# ncu evidence of one training step (run under gpurun; outputs under gpurun_out/, summaries are copied to profiles/ by
# tools/summarize_ncu.py).  B200_GRAPHS=0: kernels are launched eagerly so that every launch is visible by name.
set -u
export B200_GRAPHS=0
OUT=gpurun_out
BENCH="python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-cudnn-ref"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1400 --csv --log-file $OUT/r02_launches.csv $BENCH > $OUT/r02_launches.out 2>&1
# (2) full captures of the hot kernels, one invocation each (ncu replays every captured launch ~40 times)
cap() {  # name regex skip count
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/r02_full_$1 $BENCH > $OUT/r02_full_$1.out 2>&1
}
cap rdb_chain     rdb_chain_kernel      5 2
cap wgrad_rdb     wgrad_rdb_kernel      4 1
cap conv_igemm256 conv_igemm256_kernel  300 6
cap conv_wgrad    conv_wgrad_kernel     60 3
cap bn            "bn_reduce_kernel|bn_apply_kernel" 150 4
cap thin          "thin_to_wide_mma_kernel|wide_to_thin_mma_kernel|thin_wgrad_mma_kernel" 20 4
cap l1            l1_loss_kernel        6 2
ls -la $OUT/*.ncu-rep
