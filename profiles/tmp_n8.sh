for sl in 1 0; do
NBEST_COMM_SLOTS=$sl python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$sl bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench17_n8_sl$sl.json 2> gpurun_out/bench17_n8.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench17_n8_sl$sl.json").read().strip().split("\n")[-1])
print("N=8 slots=$sl", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"])
print({k:v["ms_per_step"] for k,v in d["kernels"].items() if k.startswith(("gemm","attn_varlen","ln","bert"))})
PY
done
