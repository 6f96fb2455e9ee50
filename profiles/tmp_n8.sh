python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_n8.json 2> gpurun_out/bench14_n8.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench14_n8.json").read().strip().split("\n")[-1])
print("N=8", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"])
for k,v in d["kernels"].items(): print(k, v["ms_per_step"], v["tflops"], v["gbs"])
PY
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_n1.json 2> gpurun_out/bench14_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench14_n1.json").read().strip().split("\n")[-1])
print("N=1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"])
PY
