B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 432 -c 8 -o gpurun_out/prof_gemm_fwd_v2 $B > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 480 -c 8 -o gpurun_out/prof_gemm_bwd_v2 $B > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:ln_bwd|ln_fwd|colsum|attn' -s 120 -c 8 -o gpurun_out/prof_misc_v2 $B > gpurun_out/ncu4.log 2>&1
ls gpurun_out | head -30
