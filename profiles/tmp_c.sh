timeout -s KILL 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout -s KILL 120 python profiles/gemm_micro.py 2>&1 | grep -E "gelu|dgrad none|bias only"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench11.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
for k,v in d["kernels"].items(): print(k, v["ms_per_step"], v["tflops"], v["gbs"])
PY
