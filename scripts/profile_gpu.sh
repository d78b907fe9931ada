#!/bin/bash
# ncu evidence for the bench step (run under gpurun, 1 GPU). Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-grid"
# launch list: device time of every launch of one timed step (skip the 3 warm-up steps' launches)
ncu --metrics gpu__time_duration.sum --clock-control none -s 5310 -c 1790 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:tr_symv_kernel -s 2600 -c 3 -o gpurun_out/prof_symv $CMD > gpurun_out/ncu_symv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prox_mgl_kernel -s 3 -c 1 -o gpurun_out/prof_prox $CMD > gpurun_out/ncu_prox.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:recon_kernel -s 3 -c 1 -o gpurun_out/prof_recon $CMD > gpurun_out/ncu_recon.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:build_w_kernel -s 3 -c 1 -o gpurun_out/prof_buildw $CMD > gpurun_out/ncu_buildw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bb_upd_kernel -s 31 -c 1 -o gpurun_out/prof_bt $CMD > gpurun_out/ncu_bt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dc_gemm_kernel -s 19 -c 1 -o gpurun_out/prof_dcgemm $CMD > gpurun_out/ncu_dcgemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tr_col_kernel -s 2600 -c 1 -o gpurun_out/prof_col $CMD > gpurun_out/ncu_col.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dc_secular_kernel -s 19 -c 1 -o gpurun_out/prof_secular $CMD > gpurun_out/ncu_secular.log 2>&1
ls -la gpurun_out
