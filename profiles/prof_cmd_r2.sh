# Round-2 captures (run under gpurun on one B200). A number printed by a run under ncu is never a bench value.
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline"
$B > gpurun_out/r2_plain.log 2>&1 || exit 1
# (1) launch list of one training step (cold-cache, serialised times: shares matter, not absolutes); 248 launches per step
ncu --metrics gpu__time_duration.sum --clock-control none -s 1488 -c 248 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_r2_l.log 2>&1
# (2) DRAM traffic of the 144 GEMM launches of one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_kernel -s 864 -c 144 --csv --log-file gpurun_out/r2_gemm_traffic.csv $B > gpurun_out/ncu_r2_t.log 2>&1
# (3) --set full: tcgen05 attention forward / backward inside the step, 8 forward-layer and 8 backward-layer GEMM launches,
#     LayerNorm forward / backward
ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 70 -c 2 -o gpurun_out/r2_prof_attn_fwd -f $B > gpurun_out/ncu_r2_af.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 70 -c 2 -o gpurun_out/r2_prof_attn_bwd -f $B > gpurun_out/ncu_r2_ab.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 888 -c 8 -o gpurun_out/r2_prof_gemm_fwd -f $B > gpurun_out/ncu_r2_gf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 936 -c 8 -o gpurun_out/r2_prof_gemm_bwd -f $B > gpurun_out/ncu_r2_gb.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:ln_fwd|ln_bwd' -s 200 -c 4 -o gpurun_out/r2_prof_ln -f $B > gpurun_out/ncu_r2_ln.log 2>&1
for f in ncu_r2_l ncu_r2_t ncu_r2_af ncu_r2_ab ncu_r2_gf ncu_r2_gb ncu_r2_ln; do tail -n 1 gpurun_out/$f.log; done
