B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:attn' -s 48 -c 4 -o gpurun_out/prof_attn_v3 $B > gpurun_out/ncu5.log 2>&1
