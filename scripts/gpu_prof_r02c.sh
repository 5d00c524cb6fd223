#!/bin/bash
# Round-2 (third session) evidence refresh: launch list of the bench command + one `ncu --set full` capture of the kernels
# that changed (warp-specialised packed lookup, tcgen05 local correlation, small-map local correlation, cluster
# normalisation, row-walking resize), raw pages as CSV.
mkdir -p gpurun_out/r02c
B="python bench.py --steps 2 --warmup 1 --workloads none --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02c/launches_bench.csv $B > gpurun_out/r02c/launches_bench.stdout 2>&1
for k in corr_lookup_packed_ws_kernel local_corr_tf32_kernel local_corr_small_kernel voxel_normalize_cluster_kernel bilinear_resize_multi_rows_kernel; do
  skip=6; [ $k = local_corr_tf32_kernel ] && skip=7      # the 80x96 level (last of the four per step)
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r02c/prof_$k -f $B > /dev/null 2>&1
  ncu -i gpurun_out/r02c/prof_$k.ncu-rep --page raw --csv > gpurun_out/r02c/prof_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out/r02c | grep -c raw.csv
