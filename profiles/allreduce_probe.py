"""NCCL all-reduce time of the gradient buckets' sizes on this box (torchrun, fp32 SUM, CUDA events, median of 20)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbest_b200.trainer import init_distributed
rank, local, world = init_distributed()
for mb in (28, 56, 95, 768):
    x = torch.ones(mb * (1 << 20) // 4, device="cuda")
    for _ in range(5): dist.all_reduce(x)
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_reduce(x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    if rank == 0:
        ms = ts[len(ts) // 2]
        print("all-reduce %4d MB fp32 on %d GPUs: %.3f ms  (algorithm bandwidth %.0f GB/s, bus bandwidth %.0f GB/s)" % (
            mb, world, ms, mb * 1.048576 / ms, mb * 1.048576 / ms * 2 * (world - 1) / world))
dist.barrier(); dist.destroy_process_group()
