#!/usr/bin/env python
"""Data-parallel parity over NCCL on real GPUs (run under torchrun, world size R >= 2; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_parity_2gpu.py

Every rank takes its slice of one global batch through DataParallelTrainer.step (bucketed NCCL all-reduce at the
attention-backward slots, per-bucket BertAdam on the side stream, --add_l2_loss on: MSE scaled by 1/R, BCE / CE summed,
per-tensor clipping after the reduce). Rank 0 then repeats the same steps on ONE GPU with the concatenated batch and
no collective, and the post-step weights must agree:  the parameter displacement after 3 optimizer steps has cosine
>= 0.999 per bucket-sized block against the single-GPU run and the same length within 2 %, the loss terms (all-reduced)
agree to 1e-3. Dropout is off (ranks draw independent masks by design). Prints one PASS / FAIL line per check. NBEST_SPARSE_EMB=1 forces
the row-sparse exchange of the word-embedding gradient (default only for XLM-R-sized tables) through the same checks.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC
    from nbest_b200.optim import BertAdam
    from nbest_b200.synth import synth_batch
    from nbest_b200.trainer import DataParallelTrainer, init_distributed
    from nbest_b200.driver import grouped_parameters

    rank, local, world = init_distributed()
    assert world >= 2, "run under torchrun with >= 2 ranks"
    dev = torch.device("cuda", local)
    hj = json.load(open(os.path.join(ROOT, "tests", "golden", "dstc2_hierarchy.json")))
    t2b = {int(k): v for k, v in hj["top2bottom"].items()}
    layers = int(os.environ.get("PARITY_LAYERS", "4"))
    per_rank = int(os.environ.get("PARITY_BATCH", "48"))
    steps = 3

    def build():
        spec = EncoderSpec.bert_base(layers=layers, hidden_dropout=0.0, attn_dropout=0.0)
        m = TOD_ASR_Transformer_STC(spec=spec, top2bottom=t2b, dropout=0.0, device=dev, none_bottoms=hj["none_bottoms"], seed=999)
        m.train()
        o = BertAdam(grouped_parameters(m, 1e-3, 2e-4), lr=1e-3, warmup=0.1, t_total=20)
        return m, o

    keys = ("ids", "seg", "trans_ids", "trans_seg", "labels")
    batches = [synth_batch("bert", 30522, type("H", (), dict(n_top=len(t2b), n_bottom=sum(len(v) for v in t2b.values()),
                                                              top2bottom=t2b, none_bottoms=hj["none_bottoms"]))(),
                           per_rank * world, 5, 128, seed=100 + s) for s in range(steps)]

    def trim(b, sl):
        d = {k: b[k][sl] for k in keys}
        S, St = int((d["ids"] > 0).sum(1).max()), int((d["trans_ids"] > 0).sum(1).max())
        out = dict(ids=d["ids"][:, :S], seg=d["seg"][:, :S], trans_ids=d["trans_ids"][:, :St], trans_seg=d["trans_seg"][:, :St],
                   labels=d["labels"])
        return {k: v.contiguous().to(dev) for k, v in out.items()}

    # ---- data-parallel run
    model, optim = build()
    init = model.flat.params.clone()
    trainer = DataParallelTrainer(model, optim, add_l2_loss=True)
    if rank == 0:
        print("INFO row-sparse embedding-gradient exchange: %s (NBEST_SPARSE_EMB=%s)" % (trainer.sparse_emb, os.environ.get("NBEST_SPARSE_EMB")))
    dp_losses = []
    for s in range(steps):
        d = trim(batches[s], slice(rank * per_rank, (rank + 1) * per_rank))
        losses = trainer.step(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"])
        dp_losses.append(trainer.global_losses(losses).clone())
    torch.cuda.synchronize()
    if rank == 0 and trainer.sparse_emb:
        print("INFO rows exchanged in the last step: %d of %d" % (trainer.last_sparse_rows, model.spec.vocab_size))
    dp_params = model.flat.params.clone()
    # every rank must hold the same replica
    chk = torch.stack([dp_params.double().sum(), dp_params.double().abs().sum()])
    gathered = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    ok = True
    if rank == 0:
        same = all(torch.equal(g, gathered[0]) for g in gathered)
        print("%s replicas identical across %d ranks after %d steps" % ("PASS" if same else "FAIL", world, steps))
        ok &= same
        # ---- single-GPU run on the concatenated batch, no collective
        ref, ropt = build()
        assert torch.equal(ref.flat.params, init)
        ref_losses = []
        for s in range(steps):
            d = trim(batches[s], slice(0, per_rank * world))
            ropt.zero_grad()
            l, _ = ref.forward_loss_backward(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], add_l2_loss=True)
            ropt.step()
            ref_losses.append(l.clone())
        torch.cuda.synchronize()
        for s in range(steps):
            # MSE: each rank's term is mean over ITS rows scaled by 1/R, the sum over ranks is the full-batch mean
            rel = float((dp_losses[s].double() - ref_losses[s].double()).abs().max() / ref_losses[s].double().abs().max())
            good = rel < 1e-3
            ok &= good
            print("%s step %d loss terms (mse, bce_final, bce_top, ce): dp %s  single %s  rel %.2e" % (
                "PASS" if good else "FAIL", s, [round(float(x), 4) for x in dp_losses[s]],
                [round(float(x), 4) for x in ref_losses[s]], rel))
        d_dp, d_ref = (dp_params - init).double(), (ref.flat.params - init).double()
        from nbest_b200.trainer import model_segments
        worst = 1.0
        for name, a, b in model_segments(model, 1):
            x, y = d_dp[a:b], d_ref[a:b]
            if float(y.norm()) == 0.0:
                continue
            cos = float(x @ y / (x.norm() * y.norm()))
            ratio = float(x.norm() / y.norm())
            good = cos >= 0.999 and abs(ratio - 1.0) < 0.02
            ok &= good
            worst = min(worst, cos)
            print("%s bucket %-8s displacement cosine %.6f  length ratio %.4f" % ("PASS" if good else "FAIL", name, cos, ratio))
        rel_w = float((dp_params.double() - ref.flat.params.double()).abs().max() / ref.flat.params.double().abs().max())
        print("%s post-step weights: worst bucket cosine %.6f, max |dp - single| / max |w| = %.2e (world %d, %d layers, %d utt/rank)" % (
            "PASS" if ok else "FAIL", worst, rel_w, world, layers, per_rank))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
