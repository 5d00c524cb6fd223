#!/bin/bash
# Quick GPU iteration: selected tests + bench + ncu launch list (+ optional full captures via $NCU_K regex list)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q ${TESTS:+-k "$TESTS"} > gpurun_out/t_quick.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list rc=$?" | tee -a gpurun_out/summary.txt
for k in $NCU_K; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/prof_$k python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu_$k rc=$?" | tee -a gpurun_out/summary.txt
done
tail -n 4 gpurun_out/t_quick.log
head -c 400 gpurun_out/bench.log; echo
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] if d.get("e2e") else None, "launches", d["gpu_launches"])
print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","avg_launch_ms","share_of_step")})
print("family", d["roofline"]["family_ms_per_step"])
print("clocks", d["clocks"])
PY
