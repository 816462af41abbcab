#!/bin/bash
# The standalone HBM-bound kernels behind utils.patches / utils.imresize at full-tile sizes: plain timing (CUDA events,
# algorithmic GB/s), then ONE launch of each under `ncu --set full`.  Usage (under gpurun): bash tools/gpu_profile_hbm.sh <tag>
set -u
TAG=${1:-hbm}
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py --one hbm_kernels > gpurun_out/${TAG}_hbm_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/${TAG}_hbm_plain.log
DSEN2_DIAG_ONCE=1 timeout 900 ncu --set full --clock-control none \
  -k regex:"extract_patches_kernel|bilinear_mirror|recompose|bicubic_tiled_kernel" -c 8 \
  -o gpurun_out/${TAG}_hbm_prof python tools/gpu_diag.py --one hbm_kernels > gpurun_out/${TAG}_hbm_ncu.log 2>&1
echo "ncu rc=$?"
