#!/bin/bash
# Run on the GPU box: parity tests (stage by stage, each under its own timeout), TF32 diagnostic, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" >> gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q > gpurun_out/t_voxel.log 2>&1; echo "voxel rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_warp.py -m gpu -x -q > gpurun_out/t_warp.log 2>&1; echo "warp rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_corr.py -m gpu -q -k "fp32 or lookup or bilinear or avg_pool or loud" > gpurun_out/t_corr_fp32.log 2>&1; echo "corr_fp32 rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python scripts/debug_tf32.py > gpurun_out/tf32_debug.log 2>&1; echo "tf32_debug rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_corr.py -m gpu -q > gpurun_out/t_corr_all.log 2>&1; echo "corr_all rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/t_voxel.log gpurun_out/t_warp.log gpurun_out/t_corr_fp32.log gpurun_out/t_corr_all.log gpurun_out/smoke.log
tail -n 25 gpurun_out/tf32_debug.log
tail -c 1500 gpurun_out/bench.log
if [ "$NCU" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list rc=$?" | tee -a gpurun_out/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_lookup -s 40 -c 2 -o gpurun_out/prof_lookup python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_lookup.log 2>&1; echo "ncu_lookup rc=$?" | tee -a gpurun_out/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_tf32 -s 3 -c 1 -o gpurun_out/prof_corr python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_corr.log 2>&1; echo "ncu_corr rc=$?" | tee -a gpurun_out/summary.txt
fi
