"""Host-side cost of one training step (Python + ctypes + tensor-map encoding + launches) vs its GPU time."""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC
from nbest_b200.optim import BertAdam
from nbest_b200.synth import synth_batch
from nbest_b200.trainer import DataParallelTrainer
hj = json.load(open("tests/golden/dstc2_hierarchy.json"))
model = TOD_ASR_Transformer_STC(spec=EncoderSpec.bert_base(), top2bottom={int(k): v for k, v in hj["top2bottom"].items()}, dropout=0.3, device="cuda")
model.train()
opt = BertAdam([dict(params=p, lr=3e-5, weight_decay=0.01) for p in model.parameters()], lr=3e-5, warmup=0.1, t_total=2300)
tr = DataParallelTrainer(model, opt)
BATCH = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = synth_batch("bert", 30522, model.hier, BATCH, 5, 128, seed=1)
d = {k: b[k].cuda() for k in ("ids", "seg", "trans_ids", "trans_seg", "labels")}
step = lambda: tr.step(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], b["lens"], b["trans_lens"])
for _ in range(3): step()
torch.cuda.synchronize()
host = []
t_all0 = time.perf_counter()
for _ in range(10):
    t0 = time.perf_counter(); step(); host.append(time.perf_counter() - t0)
torch.cuda.synchronize()
t_all = (time.perf_counter() - t_all0) / 10
from nbest_b200 import _lib
ctx = _lib.context(0)
l0 = ctx.launches(); step(); torch.cuda.synchronize(); nl = ctx.launches() - l0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
print("batch %d: host enqueue time per step: %.2f ms (min %.2f)   wall per step: %.2f ms   kernels launched per step: %d   "
      "TMA descriptor cache hits so far: %d" % (BATCH, 1e3 * sum(host) / 10, 1e3 * min(host), 1e3 * t_all, nl, ctx.tmap_cache_hits()))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(10)
