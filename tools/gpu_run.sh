#!/bin/bash
# GPU run helper (usage: tools/gpu_run.sh TAG [tests|bench|ncu WORKLOADS...])
TAG=$1; shift
KREG='regex:sample|collect|probe|perclass|col_problem|merge|emit|global|rowmax|fill_|topk|fused|sigmoid_k|decode_k|effnms|coco'
for what in "$@"; do
  case $what in
    tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log; tail -4 gpurun_out/${TAG}_tests.log;;
    newtests) python -m pytest tests/test_gpu_global.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -30 > gpurun_out/${TAG}_newtests.log; tail -12 gpurun_out/${TAG}_newtests.log;;
    bench) python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 800 gpurun_out/${TAG}_bench.err;;
    quick:*) w=${what#quick:}; python bench.py --workload ${w%%:*} $( [[ $w == *:* ]] && echo --logits ${w#*:} ) --quick --steps 20 --warmup 3 > gpurun_out/${TAG}_quick_${w/:/_}.json 2> gpurun_out/${TAG}_quick_${w/:/_}.err; tail -c 400 gpurun_out/${TAG}_quick_${w/:/_}.err; python tools/show_bench.py gpurun_out/${TAG}_quick_${w/:/_}.json;;
    ncu:*) w=${what#ncu:}; ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 150 --csv --log-file gpurun_out/${TAG}_launches_${w/:/_}.csv python bench.py --workload ${w%%:*} $( [[ $w == *:* ]] && echo --logits ${w#*:} ) --quick --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_${w/:/_}.log 2>&1; python tools/launch_summary.py gpurun_out/${TAG}_launches_${w/:/_}.csv;;
    full:*) spec=${what#full:}; w=${spec%%:*}; kn=${spec#*:}; ncu --set full --import-source on --clock-control none -k "regex:$kn" -c 1 -f -o gpurun_out/${TAG}_full_${w}_${kn} python bench.py --workload $w --quick --steps 2 --warmup 3 > gpurun_out/${TAG}_fullncu_${w}_${kn}.log 2>&1; ls -la gpurun_out/${TAG}_full_${w}_${kn}.ncu-rep;;
    globaltests) python -m pytest tests/test_gpu_global.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_globaltests.log; tail -12 gpurun_out/${TAG}_globaltests.log;;
  esac
done
