timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --tb=short -x -k "two_layer" 2>&1 | tail -15 | cut -c1-700
P='import json,sys
d=json.loads(sys.stdin.read()); k=d["roofline"]["per_kernel_ms"]
print(d["value"], d["ms_per_step"], d["clocks"]["reasons"], k)'
for v in 0 1; do for w in cfg3 cfg4; do echo $w MC=$v; VAD_LSTM_MC=$v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w 2>&1 | tail -1 | python -c "$P"; done; done
