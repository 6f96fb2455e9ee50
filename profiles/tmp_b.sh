timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench8.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"])
for k,v in d["kernels"].items(): print(k, v["ms_per_step"], v["tflops"], v["gbs"])
PY
