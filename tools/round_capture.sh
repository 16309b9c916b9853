#!/bin/bash
# One gpurun call that produces the evidence of a round: smoke, the GPU parity tests, the bench lines (ours + CPU arm),
# the ncu launch list / DRAM traffic of one forward (serial order) and of one full-pipeline batch.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/round_capture.sh [tag]'
cd "$(dirname "$0")/.."
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; head -c 600 $OUT/${TAG}_bench_line.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref_line.json 2>> $OUT/${TAG}_bench.err
echo "ref rc=$?"; head -c 300 $OUT/${TAG}_bench_ref_line.json; echo
if [ "$2" != "noncu" ]; then
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
  CTD_OVERLAP=0 timeout 600 ncu --metrics $M --clock-control none -c 600 --csv --log-file $OUT/${TAG}_traffic_bs16.csv \
    python tools/profile_forward.py 16 2 > $OUT/${TAG}_op_table.log 2>&1
  echo "ncu forward rc=$?"
  CTD_OVERLAP=0 timeout 600 ncu --metrics $M --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_traffic_pipeline_bs16.csv \
    python tools/profile_pipeline.py 16 2 > $OUT/${TAG}_pipeline.log 2>&1
  echo "ncu pipeline rc=$?"
fi
ls -la $OUT | tail -20
