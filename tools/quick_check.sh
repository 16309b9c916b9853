#!/bin/bash
# Short GPU call while iterating on kernels: the tests named on the command line first (fail fast), then a bench line
# without the extras and the ncu list of one full-pipeline batch.  Usage: gpurun -- 'bash tools/quick_check.sh TAG tests...'
cd "$(dirname "$0")/.."
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest "$@" -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
rc=$?
echo "pytest rc=$rc"; tail -15 $OUT/${TAG}_pytest.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --sustain-steps 0 > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_line.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["net_only"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_step"], d["stage_ms"])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
CTD_OVERLAP=0 timeout 600 ncu --metrics $M --clock-control none -c 1400 --csv --log-file $OUT/${TAG}_traffic_pipeline_bs16.csv \
  python tools/profile_pipeline.py 16 2 > $OUT/${TAG}_pipeline.log 2>&1
echo "ncu pipeline rc=$?"
python tools/traffic_report.py $OUT/${TAG}_traffic_pipeline_bs16.csv 2 | head -40
