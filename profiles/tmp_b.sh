timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench5.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
for k,v in d["kernels"].items(): print(k, v["ms_per_step"], v["tflops"], v["gbs"])
PY
