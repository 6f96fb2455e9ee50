python profiles/run_configs.py 2> gpurun_out/run_configs.err | tee gpurun_out/run_configs_v2.jsonl
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-transcript > gpurun_out/bench18_skip.json 2> gpurun_out/bench18.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench18_skip.json").read().strip().split("\n")[-1])
print("skip-transcript", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["streams"])
PY
