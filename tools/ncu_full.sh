#!/bin/bash
# ncu --set full --import-source on for a few kernels of the second batch / forward.  gpurun -- 'bash tools/ncu_full.sh TAG'
cd "$(dirname "$0")/.."
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
N="ncu --set full --import-source on --clock-control none"
export CTD_OVERLAP=0
timeout 300 $N -k regex:k_label_local --launch-skip 5 --launch-count 2 -f -o $OUT/${TAG}_label python tools/profile_pipeline.py 16 2 > $OUT/${TAG}_ncu1.log 2>&1
echo "label rc=$?"
timeout 300 $N -k regex:"k_phase0|k_flat2_macc" --launch-skip 6 --launch-count 2 -f -o $OUT/${TAG}_phase0 python tools/profile_pipeline.py 16 2 > $OUT/${TAG}_ncu2.log 2>&1
echo "phase0 rc=$?"
timeout 300 $N -k regex:"conv_bneck|conv_halo_kernel" --launch-skip 13 --launch-count 8 -f -o $OUT/${TAG}_conv python tools/profile_forward.py 16 2 > $OUT/${TAG}_ncu3.log 2>&1
echo "conv rc=$?"
ls -la $OUT/*.ncu-rep
