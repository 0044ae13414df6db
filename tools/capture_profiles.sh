#!/bin/bash
# Round-end evidence: launch lists and one `ncu --set full` capture per hot kernel, each only
# after the same command has exited 0 without ncu.  Run under gpurun; outputs in gpurun_out/.
#   tools/capture_profiles.sh <tag>
set -u
out=gpurun_out
tag=${1:-r2}
for st in em_real dtw convert; do
  timeout 300 python tools/prof_stage.py $st --reps 2 > $out/${tag}_${st}_plain.log 2>&1 || { echo "plain $st failed"; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $out/${tag}_launches_${st}.csv python tools/prof_stage.py $st --reps 1 > $out/${tag}_ncu_${st}.log 2>&1
done
cap() {  # stage kernel-regex skip name
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:$2" \
      --launch-skip $3 --launch-count 1 -o $out/${tag}_$4 -f python tools/prof_stage.py $1 --reps 2 \
      > $out/${tag}_ncu_$4.log 2>&1
}
cap em_real estep_tc_kernel 13 estep_tc
cap em_real mstats_tc2_kernel 14 mstats_tc2
cap dtw dtw_dist_kernel 9 dtw_dist
cap dtw dtw_dpw_kernel 9 dtw_dpw
cap convert estep_tc_kernel 1 convert_estep
ls -la $out | grep ${tag}_
