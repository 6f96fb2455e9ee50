"""CUDA-graph replay of the training step (nbest_b200.graph) against the eager step it captures.

 * replay == eager: a twin model stepped eagerly with the same by-value seeds and the same device-side step state
   (dropout salt, BertAdam schedule multiplier) follows the graphed model step for step, dropout ON, a warm-up schedule
   whose multiplier changes every step (a graph that baked the capture-time value in would diverge at once), new token
   ids and labels in the static input buffers at every replay;
 * filler sequences (graph.add_fillers, forward_loss_backward n_real) change nothing: losses and every gradient tensor of
   a batch with fillers equal those of the plain batch; batches of different token counts replay ONE graph.
"""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_add_fillers_host_properties():
    from nbest_b200.graph import add_fillers
    rng = np.random.RandomState(0)
    for trial in range(50):
        B, S = int(rng.randint(1, 40)), int(rng.randint(4, 130))
        lens = [int(x) for x in rng.randint(1, S + 1, size=B)]
        ids = torch.zeros((B, S), dtype=torch.int64)
        for b, L in enumerate(lens):
            ids[b, :L] = torch.from_numpy(rng.randint(5, 1000, size=L))
        seg = (ids > 500).long()
        n_fill, mult = (3, 256) if trial % 2 else (5, 512)
        out, seg_out, lens_out = add_fillers(ids, seg, lens, n_fill, mult, width=128)
        assert out.shape[0] == B + n_fill and out.shape[1] >= max(S, 128) and seg_out.shape == out.shape
        assert sum(lens_out) % mult == 0 and 0 < sum(lens_out) - sum(lens) < mult + n_fill
        assert torch.equal(out[:B, :S], ids) and int(out[:B, S:].sum()) == 0 and torch.equal(seg_out[:B, :S], seg)
        assert ((out > 0).sum(1).tolist() == lens_out) and all(1 <= x <= 128 for x in lens_out[B:])
        assert int(seg_out[B:].sum()) == 0
    with pytest.raises(ValueError):
        add_fillers(ids, None, lens, 1, 256)          # one filler of <= 128 tokens cannot close a 256-token gap


def _setup(dropout, layers=2, seed=5, t_total=12):
    from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC
    from nbest_b200.optim import BertAdam
    from nbest_b200.trainer import DataParallelTrainer
    hj = json.load(open(os.path.join(GOLD, "dstc2_hierarchy.json")))
    t2b = {int(k): v for k, v in hj["top2bottom"].items()}
    kw = {} if dropout else dict(hidden_dropout=0.0, attn_dropout=0.0)
    model = TOD_ASR_Transformer_STC(spec=EncoderSpec.bert_base(layers=layers, **kw), top2bottom=t2b,
                                    dropout=0.3 if dropout else 0.0, device="cuda", none_bottoms=hj["none_bottoms"], seed=seed)
    model.train()
    groups = [dict(params=p, weight_decay=0.0 if ("bias" in n or "LayerNorm" in n) else 0.01, lr=1e-3)
              for n, p in model.named_parameters()]
    opt = BertAdam(groups, lr=1e-3, warmup=0.25, t_total=t_total)
    return model, opt, DataParallelTrainer(model, opt, add_l2_loss=True)


def _batch(model, B, seed, relabel=None):
    from nbest_b200.synth import synth_batch
    b = synth_batch("bert", model.spec.vocab_size, model.hier, B, 5, 128, seed=seed)
    if relabel is not None:                      # same shapes and lengths, different token ids and labels
        g = torch.Generator().manual_seed(relabel)
        for k in ("ids", "trans_ids"):
            rnd = torch.randint(1000, 20000, b[k].shape, generator=g)
            keep = (b[k] == 0) | (b[k] == 101) | (b[k] == 102)
            b[k] = torch.where(keep, b[k], rnd)
        b["labels"] = b["labels"][torch.randperm(B, generator=g)]
    return b


@pytest.mark.gpu
def test_graph_replay_follows_the_eager_step():
    from nbest_b200 import _lib
    from nbest_b200.graph import GraphedTrainer
    mg, og, tg = _setup(dropout=True)
    me, oe, te = _setup(dropout=True)
    gt = GraphedTrainer(tg)
    ctx = _lib.context(0)
    keys = ("ids", "labels", "trans_ids", "seg", "trans_seg")
    seed_at_capture = None
    for i in range(6):
        b = _batch(mg, 24, seed=3, relabel=100 + i)
        d = {k: b[k].cuda() for k in keys}
        pinned = {k: b[k].pin_memory() for k in keys}             # replay copies host -> static inputs
        # the twin starts every step from the graphed model's state: each step is then a one-step comparison (BertAdam's
        # normalised update turns the fp32 summation-order noise of near-zero gradients — the analytically zero
        # key.bias gradients above all — into a random walk that no tolerance survives over several steps)
        with torch.no_grad():
            me.flat.params.copy_(mg.flat.params)
            if og.flat.m is not None:
                oe.flat.ensure_moments()
                oe.flat.m.copy_(og.flat.m)
                oe.flat.v.copy_(og.flat.v)
        salt, sched = gt._salt(), gt._sched()
        if i == 1:
            seed_at_capture = mg._step_seed                        # the by-value seeds the graph bakes in
        lg = gt.step(pinned["ids"], pinned["labels"], pinned["trans_ids"], pinned["seg"], pinned["trans_seg"], b["lens"],
                     b["trans_lens"]).clone()
        if i >= 1:      # twin: same by-value seeds as the captured launches + the same device-side step state
            me._step_seed = seed_at_capture
            ctx.set_step_state(salt, sched, 1.0, 1.0, torch.cuda.current_stream().cuda_stream)
        le = te.step(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], b["lens"], b["trans_lens"])
        ctx.set_step_state(0, 1.0, 1.0, 1.0, torch.cuda.current_stream().cuda_stream)
        assert torch.isfinite(lg).all()
        assert torch.allclose(lg, le, rtol=1e-5, atol=1e-5), (i, lg.tolist(), le.tolist())
        assert i == 0 or mg._step_seed == seed_at_capture + i
        moved = float((mg.flat.params - me.flat.params).abs().max())
        assert moved < 2e-6, (i, moved)                            # (observed: 1.2e-7, one ulp of a LayerNorm weight)
        assert float((og.flat.m - oe.flat.m).abs().max()) < 1e-6 and float((og.flat.v - oe.flat.v).abs().max()) < 1e-7
        if i == 1:
            first = lg
        if i == 2:      # different salt, schedule value and inputs at every replay: nothing of the capture step is replayed
            assert not torch.allclose(lg, first, rtol=1e-3)
    assert gt.captures == 1 and gt.replays == 5 and gt.eager_steps == 1 and gt.capture_error is None
    assert og._steps == oe._steps


@pytest.mark.gpu
def test_fillers_change_nothing_and_share_one_graph():
    from nbest_b200.graph import GraphedTrainer, add_fillers
    mg, og, tg = _setup(dropout=False, t_total=-1)
    me, oe, te = _setup(dropout=False, t_total=-1)
    keys = ("ids", "labels", "trans_ids", "seg", "trans_seg")
    # (1) one batch, gradients with and without fillers
    b = _batch(mg, 16, seed=11)
    d = {k: b[k].cuda() for k in keys}
    l0, _ = me.forward_loss_backward(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], add_l2_loss=True,
                                     input_lens=b["lens"], trans_input_lens=b["trans_lens"])
    g0 = me.flat.grads.clone()
    ids_f, seg_f, lens_f = add_fillers(d["ids"], d["seg"], b["lens"], 5, 512, 128)
    tid_f, tseg_f, tlens_f = add_fillers(d["trans_ids"], d["trans_seg"], b["trans_lens"], 5, 512, 128)
    assert sum(lens_f) % 512 == 0 and sum(tlens_f) % 512 == 0
    l1, head = mg.forward_loss_backward(ids_f, d["labels"], tid_f, seg_f, tseg_f, add_l2_loss=True, input_lens=lens_f,
                                        trans_input_lens=tlens_f, n_real=16)
    g1 = mg.flat.grads.clone()
    assert head.n_real == 16 and head.decode.shape[0] == 21
    # not bit-equal: the fillers change how sequences pack into 128-row attention tiles (the tile kernels' row-max bound
    # looks at the whole tile) and where the wgrad GEMMs split their token dimension — rounding-level effects. A filler
    # that leaked into the loss would move the BCE terms by ~ 5 / 21 of their value.
    assert torch.allclose(l0, l1, rtol=2e-4, atol=5e-5), (l0.tolist(), l1.tolist())
    assert float((g0 - g1).abs().max()) <= 2e-2 * float(g0.abs().max())
    cos = torch.nn.functional.cosine_similarity(g0.double(), g1.double(), dim=0)
    assert float(cos) > 0.9999, float(cos)
    me.zero_grad()
    mg.zero_grad()
    # (2) batches of different token counts fall into one bucketed shape and replay one graph
    gt = GraphedTrainer(tg, bucket=(5, 512), width=128)
    shapes = set()
    for i in range(5):
        b = _batch(mg, 16, seed=40 + i)
        shapes.add((sum(b["lens"]), sum(b["trans_lens"])))
        d = {k: b[k].cuda() for k in keys}
        with torch.no_grad():                      # one-step comparisons (see the test above)
            me.flat.params.copy_(mg.flat.params)
            if og.flat.m is not None:
                oe.flat.ensure_moments()
                oe.flat.m.copy_(og.flat.m)
                oe.flat.v.copy_(og.flat.v)
        start = mg.flat.params.clone()
        lg = gt.step(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], b["lens"], b["trans_lens"]).clone()
        le = te.step(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], b["lens"], b["trans_lens"])
        assert torch.allclose(lg, le, rtol=1e-3, atol=1e-3), (i, lg.tolist(), le.tolist())
        # the step both models took from the same state: BertAdam's normalised update amplifies rounding-level gradient
        # differences of noise-dominated tensors, so the comparison is on the whole update, not per element
        dg, de = (mg.flat.params - start).double(), (me.flat.params - start).double()
        rel = float((dg - de).norm() / de.norm())
        assert rel < 0.15, (i, rel)
        assert gt.last_head.decode.shape[0] == 21
        assert int((gt.last_head.decode[:16] != te.last_head.decode).sum()) <= 2      # (a score within rounding of 0.5 may flip)
    assert len(shapes) == 5
    assert gt.captures == 1 and gt.replays == 4 and gt.eager_steps == 1 and gt.capture_error is None, \
        (gt.captures, gt.replays, gt.capture_error)
