for k in 1 2 4 12; do
export NBEST_BUCKET_LAYERS=$k
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$k bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8_k$k.json 2> gpurun_out/bench_n8.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n8_k$k.json").read().strip().split("\n")[-1])
print("k=$k", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done
NBEST_KEEP_NCCL_DEBUG=1 NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -iE "nvls|algo|channels|Connected" | sort | uniq -c | head -12
