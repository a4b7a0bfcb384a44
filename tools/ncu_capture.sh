#!/bin/bash
# One `ncu --set full` capture per hot kernel (first launch of each) of one C2 motion_correct step, plus the launch
# list (device time of every launch) of the same command.  Usage (on the GPU box): bash tools/ncu_capture.sh <tag>
set -u
tag=${1:-r01}
mkdir -p gpurun_out
python tools/profile_step.py --iterations 100 > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${tag}.csv python tools/profile_step.py --iterations 100 > gpurun_out/ncu_launches.log 2>&1
i=0
for k in 'local_loss_tile_kernel' 'local_coefficient_kernel' 'rows_forward_poly<.int.1' 'rows_forward_poly<.int.2' 'cols_forward_p2<.int.1024>' \
         'rows_inverse_argmax_poly' 'cols_inverse_p2<.int.1024>' 'rows_forward_real2n' 'cols_forward_p2<.int.4096>' \
         'cols_inverse_p2<.int.4096>' 'rows_inverse_argmax_real2n' \
         'warp_tma_kernel' 'lattice_xinterp_kernel' 'xc_leave_one_out_kernel' 'stats_partial_kernel'; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k regex:"$k" -c 1 -o gpurun_out/prof_${tag}_$i python tools/profile_step.py --iterations 3 > gpurun_out/ncu_$i.log 2>&1
  echo "$k -> $(tail -1 gpurun_out/ncu_$i.log)"
  # keep what travels back small: raw metrics as csv, drop the .ncu-rep
  if [ -f gpurun_out/prof_${tag}_$i.ncu-rep ]; then
    ncu -i gpurun_out/prof_${tag}_$i.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_${i}_raw.csv 2>/dev/null
    rm -f gpurun_out/prof_${tag}_$i.ncu-rep
  fi
done
du -sh gpurun_out
