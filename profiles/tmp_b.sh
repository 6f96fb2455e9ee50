for cg in 2 1; do
NBEST_GEMM_CTA_GROUP=$cg python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench10_cg$cg.json 2> gpurun_out/bench10.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench10_cg$cg.json").read().strip().split("\n")[-1])
print("CG=$cg", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
for k,v in d["kernels"].items(): print(k, v["ms_per_step"], v["tflops"], v["gbs"])
PY
done
