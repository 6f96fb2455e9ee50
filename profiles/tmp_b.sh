timeout -s KILL 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench9.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"])
for k,v in d["kernels"].items(): print(k, v["ms_per_step"], v["tflops"], v["gbs"])
PY
