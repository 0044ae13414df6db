#!/bin/bash
# Round-end evidence: launch lists and one `ncu --set full` capture per hot kernel, each only
# after the same command has exited 0 without ncu.  Run under gpurun; outputs in gpurun_out/.
set -u
out=gpurun_out
tag=${1:-r1b}
for st in em dtw convert; do
  timeout 200 python tools/prof_stage.py $st --reps 2 > $out/${tag}_${st}_plain.log 2>&1 || { echo "plain $st failed"; continue; }
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $out/${tag}_launches_${st}.csv python tools/prof_stage.py $st --reps 2 > $out/${tag}_ncu_${st}.log 2>&1
done
cap() {  # stage kernel-regex skip name
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name "regex:$2" \
      --launch-skip $3 --launch-count 1 -o $out/${tag}_$4 -f python tools/prof_stage.py $1 --reps 2 \
      > $out/${tag}_ncu_$4.log 2>&1
}
cap em estep_tc_kernel 1 estep_tc
cap em mstats_tc_kernel 2 mstats_tc
cap em gmm_finalize_kernel 2 finalize
cap dtw dtw_dist_kernel 9 dtw_dist
cap dtw dtw_dpw_kernel 9 dtw_dpw
cap convert estep_tc_kernel 1 convert_estep
cap convert convert_condmean_kernel 1 convert_condmean
cap convert convert_mlpg_kernel 1 convert_mlpg
ls -la $out | grep ${tag}_
