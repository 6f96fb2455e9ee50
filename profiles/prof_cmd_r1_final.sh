# Round-1 final captures (run under gpurun on one B200). A number printed by a run under ncu is never a bench value.
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_final.log 2>&1 || exit 1
# (1) launch list of one training step (cold-cache, serialised times: shares matter, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 270 --csv --log-file gpurun_out/launches_r1_final.csv $B > gpurun_out/ncu_l.log 2>&1
# (2) DRAM traffic of the 144 GEMM launches of one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_kernel -s 1008 -c 144 --csv --log-file gpurun_out/gemm_traffic_r1_final.csv $B > gpurun_out/ncu_t.log 2>&1
# (3) --set full of the dominant kernels: 8 consecutive GEMM launches of a forward layer, 8 of a backward layer, and the
#     attention / LayerNorm / column-sum / BertAdam kernels
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1032 -c 8 -o gpurun_out/prof_gemm_fwd_final -f $B > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1080 -c 8 -o gpurun_out/prof_gemm_bwd_final -f $B > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:attn_fwd|ln_fwd' -s 200 -c 4 -o gpurun_out/prof_misc_fwd_final -f $B > gpurun_out/ncu_m.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:attn_bwd|ln_bwd|colsum|adam_' -s 400 -c 10 -o gpurun_out/prof_misc_bwd_final -f $B > gpurun_out/ncu_m2.log 2>&1
for f in ncu_l ncu_t ncu_f ncu_b ncu_m ncu_m2; do tail -n 1 gpurun_out/$f.log; done
