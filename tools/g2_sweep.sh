#!/bin/bash
# G2 MSM at 2^20: window widths without a table, table widths, and the batched-affine pre-reduction forced on
# (MSM_AFFINE = option msm_affine).  usage: tools/g2_sweep.sh [log_n]
L=${1:-20}
echo "# plain"; MSM_G2=1 python tools/tune_msm.py $L 13 17 | cut -c1-200
for c in 14 16 18 20; do echo "# table c=$c"; for a in 1 2 3; do echo "## msm_affine=$a"; MSM_AFFINE=$a MSM_G2=1 MSM_TABLE=$c python tools/tune_msm.py $L $c $c | cut -c1-200; done; done
