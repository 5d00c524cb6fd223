#!/bin/bash
# Round-2 evidence: the launch list of the bench command and one `ncu --set full` capture per dominant kernel.
mkdir -p gpurun_out/r02
B="python bench.py --steps 2 --warmup 1 --workloads none --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_bench.csv $B > gpurun_out/r02/launches_bench.stdout 2>&1
for k in corr_lookup_packed_kernel corr_tf32_pair_kernel voxel_vote_atomic_kernel local_corr_vec_kernel pool_pyramid_packed_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/r02/prof_$k $B > /dev/null 2>&1
done
# HREM: voxel kernels (order-free K1 and the exact tile-binned path)
BENCH_VOXEL_CASES="x1 uniform" ncu --set full --clock-control none --import-source on -k regex:voxel_bin_kernel -s 1 -c 1 -o gpurun_out/r02/prof_voxel_bin_kernel python scripts/bench_voxel.py > /dev/null 2>&1
BENCH_VOXEL_CASES="x1 uniform" ncu --set full --clock-control none --import-source on -k regex:voxel_plane_kernel -s 1 -c 1 -o gpurun_out/r02/prof_voxel_plane_kernel python scripts/bench_voxel.py > /dev/null 2>&1
BENCH_VOXEL_CASES="x1 uniform" ncu --set full --clock-control none --import-source on -k regex:voxel_vote_atomic_kernel -s 1 -c 1 -o gpurun_out/r02/prof_voxel_vote_atomic_hrem python scripts/bench_voxel.py > /dev/null 2>&1
ls -la gpurun_out/r02
