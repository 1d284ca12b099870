#!/bin/bash
# Round-2 final measurement campaign on one B200 (run under gpurun): GPU test log, bench lines for cfg1-4, the reference
# arm; with a second argument "ncu" also ncu --set full captures of one cfg2 and one cfg4 step.  Outputs under
# gpurun_out/ with the given tag.
TAG=${1:-r02f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -1 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
python bench.py --gpus 1 --steps 50 --warmup 5 --no-extras --no-cpu-baseline --no-gpu-library-baseline > gpurun_out/${TAG}_bench_cfg2_50steps.json 2>> gpurun_out/${TAG}_bench_cfg2.err
for W in cfg1 cfg3 cfg4; do
  python bench.py --gpus 1 --steps 20 --warmup 5 --workload $W > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
done
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for f in gpurun_out/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
if [ "$2" = "ncu" ]; then  # (two reports of one step each: gpurun brings back at most 64 MiB)
timeout 900 ncu --set full --clock-control none -k "regex:conv|finalize|enc1_fused" -s 45 -c 16 -f -o gpurun_out/${TAG}_prof_cfg2 \
  python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-gpu-library-baseline > gpurun_out/${TAG}_ncu_cfg2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_cfg2.log | cut -c1-200
timeout 900 ncu --set full --clock-control none -k "regex:conv|finalize|enc1_fused" -s 27 -c 10 -f -o gpurun_out/${TAG}_prof_cfg4 \
  python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-gpu-library-baseline --workload cfg4 > gpurun_out/${TAG}_ncu_cfg4.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_cfg4.log | cut -c1-200
fi
