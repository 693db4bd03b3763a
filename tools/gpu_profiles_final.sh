#!/bin/bash
TAG=r02c
KREG='regex:sample|collect|probe|perclass|col_problem|merge|emit|global|rowmax|fill_|topk|fused|sigmoid_k|decode_k|effnms|coco'
for w in c2 c3 c4 c5 c2:clustered; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 120 --csv \
    --log-file gpurun_out/${TAG}_launches_${w/:/_}.csv python bench.py --workload ${w%%:*} \
    $( [[ $w == *:* ]] && echo --logits ${w#*:} ) --quick --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_${w/:/_}.log 2>&1
  echo "== $w"; python tools/launch_summary.py gpurun_out/${TAG}_launches_${w/:/_}.csv | tail -n +2
done > gpurun_out/${TAG}_launch_summary.txt 2>&1
full() {
  timeout 300 ncu --set full --import-source on --clock-control none -k "regex:$2" -c 1 -f \
    -o gpurun_out/${TAG}_full_${1/:/_}_$3 python bench.py --workload ${1%%:*} $( [[ $1 == *:* ]] && echo --logits ${1#*:} ) \
    --quick --steps 2 --warmup 3 > gpurun_out/${TAG}_fullncu_$3.log 2>&1
  python tools/ncu_summary.py gpurun_out/${TAG}_full_${1/:/_}_$3.ncu-rep >> gpurun_out/${TAG}_ncu_summary.txt 2>&1
  python tools/ncu_lines.py gpurun_out/${TAG}_full_${1/:/_}_$3.ncu-rep "$2" 25 > gpurun_out/${TAG}_ncu_lines_$3.txt 2>&1
  rm -f gpurun_out/${TAG}_full_${1/:/_}_$3.ncu-rep
}
rm -f gpurun_out/${TAG}_ncu_summary.txt
full c2 collect_cols4_kernel collect_cols4
full c2 'col_problem_kernel' col_problem_finish
full c2:clustered 'col_problem_kernel' col_problem_finish_clustered
full c2 probe_warp_kernel probe_warp
cat gpurun_out/${TAG}_launch_summary.txt
