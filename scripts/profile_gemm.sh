#!/bin/bash
# ncu --set full captures of the FP64 tensor-core kernels of the eigensolver's stages 2/3 and of the upper-triangle prox
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
CMD="python scripts/profile_step.py 1"
for spec in "bb_upd_kernel 3" "bb_y_kernel 3" "dc_gemm_kernel 4" "prox_mgl_upper_kernel 0" "dc_secular_kernel 4" "dc_prepare_kernel 4"; do
  set -- $spec
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/r02_prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
done
ls -la gpurun_out | tail -8
