"""CPU: the oracle restatement against the golden vectors the LIVE reference produced (oracle/make_golden.py).

These tests pin the checker itself: if oracle/stc_oracle.py drifts from the reference's algorithm, they fail here,
on CPU, before any GPU parity claim is made against it."""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hier():
    from oracle import stc_oracle as O
    hj = json.load(open(os.path.join(GOLD, "dstc2_hierarchy.json")))
    return O.Hierarchy({int(k): v for k, v in hj["top2bottom"].items()}, hj["none_bottoms"]), hj


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", ["bert_l2_small", "bert_nol2_small", "xlmr_l2_small"])
def test_oracle_reproduces_reference_vectors(name):
    from oracle import stc_oracle as O
    torch.set_num_threads(min(8, os.cpu_count()))
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.EncoderConfig(**json.loads(str(fx["cfg"])))
    hier, _ = _hier()
    hp = json.loads(str(fx["hp"]))
    params = O.init_params(cfg, hier, seed=int(fx["seed"]), style="perturbed")
    chk = float(sum(v.double().abs().sum() for v in params.values()))
    assert abs(chk - float(fx["weight_checksum"][0])) / chk < 1e-9
    batch = {k: torch.from_numpy(fx[k]) for k in ("ids", "seg", "trans_ids", "trans_seg", "labels")}
    state = {}
    names = [str(x) for x in fx["grad_names"]]
    for step in range(int(fx["n_steps"])):
        terms, grads, (top, bottoms, final, asr, trans) = O.train_step(params, cfg, hier, batch, state, hp)
        assert _rel(top.detach(), fx["top_%d" % step]) < 1e-4
        assert _rel(final.detach(), fx["final_%d" % step]) < 1e-4
        assert _rel(asr.detach(), fx["asr_%d" % step]) < 1e-4
        assert _rel(trans.detach(), fx["trans_%d" % step]) < 1e-4
        assert abs(terms["total"] - float(fx["total_%d" % step][0])) / abs(float(fx["total_%d" % step][0])) < 1e-5
        rec = sum(terms[k] for k in ("mse", "bce_final", "bce_top", "ce") if k in terms) / batch["ids"].shape[0]
        assert abs(rec - float(fx["rec_%d" % step][0])) / abs(float(fx["rec_%d" % step][0])) < 1e-5    # loss_record
        for n, nr in zip(names, fx["gradnorm_%d" % step]):
            if grads[n] is None:
                assert nr == 0.0 and "pooler" in n
            elif "attention.self.key.bias" not in n:
                assert abs(float(grads[n].double().norm()) - nr) <= 1e-3 * nr + 1e-9, n
        for k in fx.files:
            if k.startswith("grad_%d_" % step):
                assert _rel(grads[k[len("grad_%d_" % step):]], fx[k]) < 2e-3, k
        if step == 0:
            assert np.array_equal(O.decode(hier, top, bottoms), fx["decode"])
    for n, nr in zip(names, fx["post_norm"]):
        assert abs(float(params[n].double().norm()) - nr) <= 1e-5 * nr, n
    for k in fx.files:
        if k.startswith("post_") and k != "post_norm":
            assert float((params[k[5:]] - torch.from_numpy(fx[k])).abs().max()) < 2e-5, k


def test_pack_batch_inverts_reference_padding():
    """The packed layout is the bit-exact un-padding of what utils/bert_xlnet_inputs.py produced (fixture)."""
    from oracle import stc_oracle as O
    fx = np.load(os.path.join(GOLD, "packing_valid24.npz"))
    ids, seg, lens = fx["ids_default"], fx["seg_default"], fx["lens_default"]
    pk = O.pack_batch(ids, seg, "bert")
    assert np.array_equal(pk["lens"], lens)
    assert pk["T"] == int(lens.sum()) and pk["cu_seqlens"][-1] == pk["T"]
    # re-pad and compare with the reference tensors
    B, S = ids.shape
    back = np.zeros((B, S), dtype=np.int64)
    back_seg = np.zeros((B, S), dtype=np.int64)
    for b in range(B):
        sl = slice(pk["cu_seqlens"][b], pk["cu_seqlens"][b + 1])
        back[b, :lens[b]] = pk["tokens"][sl]
        back_seg[b, :lens[b]] = pk["seg"][sl]
        assert np.array_equal(pk["pos"][sl], np.arange(lens[b]))
        assert (pk["seq_of"][sl] == b).all() and pk["key_valid"][sl].all()
    assert np.array_equal(back, ids) and np.array_equal(back_seg, seg)
    # layout facts of A.1: [CLS] first, segment 0 then 1 from the first [SEP], last token [SEP]
    for b in range(B):
        row, sg = ids[b, :lens[b]], seg[b, :lens[b]]
        first_sep = int(np.argmax(row == 102))
        assert row[0] == 101 and row[-1] == 102
        assert (sg[:first_sep] == 0).all() and (sg[first_sep:] == 1).all()


def test_pack_batch_xlmr_quirks_and_edge_cases():
    from oracle import stc_oracle as O
    ids = np.array([[0, 9, 8, 2, 1, 1], [0, 5, 1, 7, 2, 1], [1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 0]], dtype=np.int64)
    pk = O.pack_batch(ids, None, "xlm-roberta")
    # <pad>=1 is > 0 so padded rows keep their full length; an all-<s> row (ids == 0) has length 0
    assert pk["lens"].tolist() == [6, 6, 6, 0]
    assert pk["key_valid"][:6].tolist() == [0, 1, 1, 1, 1, 1]                  # <s>=0 masked as a key, <pad> is not
    assert pk["pos"][:6].tolist() == [2, 3, 4, 5, 1, 1]                        # cumsum(ids != 1) * (ids != 1) + 1
    assert pk["pos"][6:12].tolist() == [2, 3, 1, 4, 5, 1]
    assert pk["pos"][12:18].tolist() == [1] * 6
    bert = O.pack_batch(np.array([[101, 7, 102, 0, 0], [0, 0, 0, 0, 0]], dtype=np.int64), None, "bert")
    assert bert["lens"].tolist() == [3, 0] and bert["T"] == 3


def test_loss_semantics_sum_reduced_and_none_target():
    """BCE/CE are sums over the batch (doubling the batch doubles them), the MSE is a mean; an empty group targets NONE."""
    from oracle import stc_oracle as O
    hier, _ = _hier()
    g = torch.Generator().manual_seed(0)
    B = 5
    f = torch.randn(B, 768, generator=g)
    params = O.init_params(O.EncoderConfig.bert_base(layers=0, vocab_size=10, max_position=4), hier, seed=3, style="perturbed")
    top, bottoms, final = O.head_forward(params, hier, f)
    labels = torch.zeros(B, hier.n_bottom)
    labels[0, hier.top2bottom[2][3]] = 1
    labels[1, 3] = 1
    t1, terms1 = O.total_loss(hier, top, bottoms, final, labels, f, f + 1.0, add_l2_loss=True)
    cat = lambda t: torch.cat([t, t])
    t2, terms2 = O.total_loss(hier, cat(top), {k: cat(v) for k, v in bottoms.items()}, cat(final), cat(labels), cat(f),
                              cat(f) + 1.0, add_l2_loss=True)
    for k in ("bce_final", "bce_top", "ce"):
        assert abs(float(terms2[k]) - 2 * float(terms1[k])) < 1e-3 * abs(float(terms1[k]))
    assert abs(float(terms2["mse"]) - float(terms1["mse"])) < 1e-6 and abs(float(terms1["mse"]) - 1.0) < 1e-6
    # CE of a group without an active label uses the last (NONE) column
    k = hier.group_tops[1]
    q = bottoms["lin_%d" % k]
    ce_k = -torch.log(q[:, -1] + 1e-12).sum()
    sub = labels[:, hier.top2bottom[k]]
    assert float(sub.sum()) == 0.0
    ces = []
    for kk in hier.group_tops:
        ids = hier.top2bottom[kk]
        s = labels[:, ids]
        tgt = torch.where(s.sum(1) == 0, torch.full((B,), len(ids) - 1), s.argmax(1))
        ces.append(-torch.log(bottoms["lin_%d" % kk][torch.arange(B), tgt] + 1e-12).sum())
    assert abs(float(sum(ces) / len(ces)) - float(terms1["ce"])) < 1e-5
    assert float(ces[1]) == pytest.approx(float(ce_k))


def test_bertadam_first_step_has_zero_lr_and_no_bias_correction():
    from oracle import stc_oracle as O
    p = {"clf.x.weight": torch.ones(4, 4)}
    g = {"clf.x.weight": torch.full((4, 4), 0.01)}
    st = {}
    O.bertadam_step(p, g, st, lr=0.1, bert_lr=0.1, warmup=0.1, t_total=10)
    assert torch.equal(p["clf.x.weight"], torch.ones(4, 4))                   # schedule(0) = 0
    assert st["clf.x.weight"]["step"] == 1
    m0 = st["clf.x.weight"]["m"].clone()
    assert torch.allclose(m0, torch.full((4, 4), 0.001))                      # m advanced although p did not move
    O.bertadam_step(p, g, st, lr=0.1, bert_lr=0.1, warmup=0.1, t_total=10)
    # step 1: progress 0.1 >= warmup -> multiplier max((0.1-1)/(0.1-1), 0) = 1; update = m/(sqrt(v)+1e-6) + 0.01*p
    m = 0.9 * 0.001 + 0.1 * 0.01
    v = 0.999 * (0.001 * 0.01 ** 2) + 0.001 * 0.01 ** 2
    expect = 1.0 - 0.1 * (m / (v ** 0.5 + 1e-6) + 0.01 * 1.0)
    assert abs(float(p["clf.x.weight"][0, 0]) - expect) < 1e-5
    # per-tensor clipping: a large gradient is scaled to unit norm before the moments
    p2, st2 = {"bert_encoder.b.bias": torch.zeros(100)}, {}
    O.bertadam_step(p2, {"bert_encoder.b.bias": torch.full((100,), 5.0)}, st2, 0.1, 0.1, 0.1, 10)
    assert torch.allclose(st2["bert_encoder.b.bias"]["m"], torch.full((100,), 0.1 * 5.0 / (50.0 + 1e-6)), rtol=1e-5)
    assert O.param_hyper("bert_encoder.x.LayerNorm.weight", 1.0, 2.0) == (2.0, 0.0)
    assert O.param_hyper("clf.top_linear_layer.weight", 1.0, 2.0) == (1.0, 0.01)
