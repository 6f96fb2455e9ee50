"""Per-bucket timeline of one data-parallel training step (torchrun, N >= 2): when the backward announces each gradient
bucket / collective slot, when each bucket's all-reduce starts and ends on the comm stream, when its BertAdam update starts
and ends on the optimizer stream — CUDA-event stamps in ms since the step started (rank 0 prints; max over ranks is not
taken: ranks are symmetric). BASELINE configs[1] batch shape (B = 256 per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/dp_timeline.py
"""
import json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC
from nbest_b200.optim import BertAdam
from nbest_b200.synth import synth_batch
from nbest_b200.trainer import DataParallelTrainer, NcclGradPool, init_distributed
rank, local, world = init_distributed()
pool = NcclGradPool("cuda:%d" % local)
hj = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/dstc2_hierarchy.json")))
with pool:
    model = TOD_ASR_Transformer_STC(spec=EncoderSpec.bert_base(), top2bottom={int(k): v for k, v in hj["top2bottom"].items()},
                                    dropout=0.3, device="cuda:%d" % local, none_bottoms=hj["none_bottoms"])
pool.register()
model.train()
opt = BertAdam([dict(params=p, lr=3e-5, weight_decay=0.01) for p in model.parameters()], lr=3e-5, warmup=0.1, t_total=2300)
tr = DataParallelTrainer(model, opt)
b = synth_batch("bert", 30522, model.hier, 256, 5, 128, seed=999 + 1000 * rank)
d = {k: b[k].cuda() for k in ("ids", "seg", "trans_ids", "trans_seg", "labels")}
step = lambda: tr.step(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], b["lens"], b["trans_lens"])
for _ in range(8):
    step()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
plain = e0.elapsed_time(e1) / 10
dist.barrier()
tr.start_timeline()
step()
tl = tr.stop_timeline()
if rank == 0:
    print("# gradient buffer in NCCL-registered memory: %s %s" % (pool.ok, pool.why or ""))
    print("# %d x B200, B = 256 per GPU; step without stamps %.3f ms; buckets: %s" % (
        world, plain, ", ".join("%s %.0f MB" % (n, (e - s) * 4 / 2**20) for n, s, e in tr.bucketer.segments)))
    print("# tag                               ms since step start")
    for tag, ms in tl:
        print("%-36s %8.3f" % (tag, ms))
    ar = {}
    for tag, ms in tl:
        k, _, n = tag.partition(":")
        if k in ("allreduce_start", "allreduce_end", "adam_start", "adam_end"):
            ar.setdefault(n, {})[k] = ms
    print("# bucket      all-reduce ms   adam ms")
    for n, v in ar.items():
        print("%-12s %8.3f %12.3f" % (n, v.get("allreduce_end", 0) - v.get("allreduce_start", 0), v.get("adam_end", 0) - v.get("adam_start", 0)))
    last_bwd = max(ms for tag, ms in tl if tag.startswith("backward:"))
    end = max(ms for tag, ms in tl)
    print("# last backward announcement at %.3f ms, step end at %.3f ms: exposed tail %.3f ms" % (last_bwd, end, end - last_bwd))
dist.barrier(); dist.destroy_process_group()
