#!/bin/bash
# ncu --set full --import-source on for the refine kernels of the second batch.  gpurun -- 'bash tools/ncu_refine.sh TAG'
cd "$(dirname "$0")/.."
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
N="ncu --set full --import-source on --clock-control none"
export CTD_OVERLAP=0
timeout 300 $N -k regex:k_label_local --launch-skip 5 --launch-count 1 -f -o $OUT/${TAG}_label python tools/profile_pipeline.py 16 2 > $OUT/${TAG}_ncu1.log 2>&1
echo "label rc=$?"
timeout 300 $N -k regex:"k_phase0|k_flat2_macc|k_mapply|k_xor" --launch-skip 13 --launch-count 4 -f -o $OUT/${TAG}_sweeps python tools/profile_pipeline.py 16 2 > $OUT/${TAG}_ncu2.log 2>&1
echo "sweeps rc=$?"
ls -la $OUT/*.ncu-rep
