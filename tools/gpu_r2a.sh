#!/bin/bash
# GPU run: tests, full bench, ncu launch lists per BASELINE config (usage: tools/gpu_r2a.sh TAG)
TAG=${1:-r2a}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
for w in c1 c3 c4 c5 c2s; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_$w.csv \
    python bench.py --workload $w --quick --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_$w.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_c2clu.csv \
  python bench.py --workload c2 --logits clustered --quick --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_c2clu.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
tail -c 600 gpurun_out/${TAG}_bench.err
