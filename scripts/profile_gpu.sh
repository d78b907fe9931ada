#!/bin/bash
# ncu evidence for the round-2 step (run under gpurun, 1 GPU).  Outputs under gpurun_out/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
CMD="python scripts/profile_step.py 4"
# launch list: device time of every launch of ONE iteration (skip the first three iterations: 3 x 1770 launches + setup)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 5330 -c 1775 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
# full captures of the kernels DESIGN.md quotes
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tr_symv_kernel -s 2100 -c 2 -o gpurun_out/r02_prof_symv $CMD > gpurun_out/ncu_symv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:recon_bulk_kernel -s 2 -c 1 -o gpurun_out/r02_prof_recon $CMD > gpurun_out/ncu_recon.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prox_mgl_kernel -s 2 -c 1 -o gpurun_out/r02_prof_prox $CMD > gpurun_out/ncu_prox.log 2>&1
GG_TR_BLOCKED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sytrd_panel_kernel -s 56 -c 1 -o gpurun_out/r02_prof_panel python scripts/profile_step.py 2 > gpurun_out/ncu_panel.log 2>&1
GG_TR_BLOCKED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sytrd_syr2k_kernel -s 56 -c 1 -o gpurun_out/r02_prof_syr2k python scripts/profile_step.py 2 > gpurun_out/ncu_syr2k.log 2>&1
ls -la gpurun_out | tail -12
