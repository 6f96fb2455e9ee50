# Round-2 final-state runs on one B200 (gpurun). A number printed by a run under ncu is never a bench value.
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gputests.log 2>&1; tail -n 2 gpurun_out/r2_final_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --batch 16 --bucket 2,64 --steps 50 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_final_b16.json 2> gpurun_out/r2_final_b16.err; echo "b16 rc=$?"
timeout 200 python profiles/host_overhead.py 256 2>&1 | head -n 1 > gpurun_out/r2_host_overhead.txt
timeout 200 python profiles/host_overhead.py 16 2>&1 | head -n 1 >> gpurun_out/r2_host_overhead.txt
cat gpurun_out/r2_host_overhead.txt
B="python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline --no-eager-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1488 -c 248 --csv --log-file gpurun_out/r2_final_launches.csv $B > gpurun_out/ncu_r2_final_l.log 2>&1; tail -n 1 gpurun_out/ncu_r2_final_l.log
