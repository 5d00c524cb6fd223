#!/bin/bash
# Round-end evidence run (one B200): full GPU test suite, smoke, default bench, HREM workloads, per-kernel tables,
# ncu launch list of the bench command and full captures of the dominant kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
V=${V:-v5}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" | tee gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench_$V.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$V.log 2>&1; echo "bench_ref rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --workload hrem_dt1 --no-cpu-baseline > gpurun_out/bench_hrem_dt1_$V.log 2>&1; echo "hrem_dt1 rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --workload hrem_dt4 --no-cpu-baseline --steps 10 > gpurun_out/bench_hrem_dt4_$V.log 2>&1; echo "hrem_dt4 rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python scripts/bench_kernels.py > gpurun_out/kernel_roofline_table_$V.md 2>&1; echo "kernels rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python scripts/bench_family.py > gpurun_out/eemflow_family_amortised_$V.log 2>&1; echo "family rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_$V.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list rc=$?" | tee -a gpurun_out/summary.txt
for k in corr_lookup_kernel corr_tf32_pair voxel_vote_atomic local_corr_vec; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_${k}_$V python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu_$k rc=$?" | tee -a gpurun_out/summary.txt
done
tail -n 3 gpurun_out/t_all.log gpurun_out/smoke.log
tail -c 600 gpurun_out/bench_$V.log
