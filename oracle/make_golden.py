"""Pin the oracle against the live reference and write golden vectors (runs ONLY in the build container).

It imports the UNMODIFIED reference modules from /root/reference (models.model, models.optimization, utils.STC_util,
utils.bert_xlnet_inputs, n_best_asr_bert's loss/decode functions) together with the installed HuggingFace BertModel /
XLMRobertaModel (the third-party dependency that carries the arithmetic), feeds them the oracle's
deterministic weights and synthetic inputs, asserts that oracle/stc_oracle.py agrees to fp32 round-off, and stores the
REFERENCE's outputs as small fixtures under tests/golden/. /root/reference does not exist on the GPU box: tests only
read the fixtures.

    python oracle/make_golden.py
"""
import json
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import stc_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.modules.setdefault("gpustat", types.ModuleType("gpustat"))           # utils/gpu_selection.py:11 (absent here)
    import transformers
    import transformers.optimization as topt
    if not hasattr(topt, "AdamW"):                                           # n_best_asr_bert.py:17 (removed in HF 5.x)
        topt.AdamW = torch.optim.AdamW
        transformers.AdamW = torch.optim.AdamW
    argv, sys.argv = sys.argv, ["x"]
    try:
        import n_best_asr_bert as nb
    finally:
        sys.argv = argv
    from models.model import make_model
    from models.optimization import BertAdam
    from utils.STC_util import reverse_top2bottom
    from utils.bert_xlnet_inputs import prepare_inputs_for_roberta
    return nb, make_model, BertAdam, reverse_top2bottom, prepare_inputs_for_roberta


def hf_encoder(cfg):
    import transformers
    if cfg.kind == "xlm-roberta":
        c = transformers.XLMRobertaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
                                          num_attention_heads=cfg.heads, intermediate_size=cfg.intermediate,
                                          max_position_embeddings=cfg.max_position, type_vocab_size=cfg.type_vocab,
                                          layer_norm_eps=cfg.ln_eps, pad_token_id=1, bos_token_id=0, eos_token_id=2)
        c._attn_implementation = "eager"
        return transformers.XLMRobertaModel(c)
    c = transformers.BertConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
                                num_attention_heads=cfg.heads, intermediate_size=cfg.intermediate,
                                max_position_embeddings=cfg.max_position, type_vocab_size=cfg.type_vocab,
                                layer_norm_eps=cfg.ln_eps)
    c._attn_implementation = "eager"
    return transformers.BertModel(c)


def export_hierarchy():
    """The real DSTC2 label hierarchy from the reference fixture -> tests/golden/dstc2_hierarchy.json."""
    m = torch.load(os.path.join(REF, "dstc2_data/processed_data/raw/memory.pt"))
    t2b = {int(k): [int(x) for x in v] for k, v in m["top2bottom_dict"].items()}
    none_b = [int(i) for i, l in m["idx2label"].items() if l.endswith("NONE")]
    d = dict(top2bottom={str(k): v for k, v in t2b.items()}, none_bottoms=none_b,
             idx2label={str(k): v for k, v in m["idx2label"].items()})
    with open(os.path.join(GOLD, "dstc2_hierarchy.json"), "w") as f:
        json.dump(d, f, indent=0)
    return O.Hierarchy(t2b, none_b), m


def synth_batch(cfg, hier, B, max_len, seed, with_trans=True):
    """Token-id level synthetic batch in the reference's padded layout."""
    rng = np.random.RandomState(seed)
    cls, sep, pad = (0, 2, 1) if cfg.kind == "xlm-roberta" else (101 % cfg.vocab_size, 102 % cfg.vocab_size, 0)
    lo = 5 if cfg.vocab_size < 2000 else 1000

    def stream(lmin, lmax):
        lens = rng.randint(lmin, lmax + 1, size=B)
        S = int(lens.max())
        ids = np.full((B, S), pad, dtype=np.int64)
        seg = np.zeros((B, S), dtype=np.int64)
        for b, L in enumerate(lens):
            row = rng.randint(lo, cfg.vocab_size, size=L)
            row[0] = cls
            nsys = rng.randint(2, max(3, L // 3))
            seps = sorted(set([nsys] + list(rng.choice(np.arange(nsys + 1, L), size=min(4, L - nsys - 1), replace=False)) + [L - 1]))
            row[seps] = sep
            ids[b, :L] = row
            seg[b, nsys:L] = 1
        return ids, seg

    ids, seg = stream(max(8, max_len // 3), max_len)
    batch = dict(ids=torch.from_numpy(ids), seg=torch.from_numpy(seg))
    if with_trans:
        tids, tseg = stream(6, max(8, max_len // 2))
        batch.update(trans_ids=torch.from_numpy(tids), trans_seg=torch.from_numpy(tseg))
    labels = np.zeros((B, hier.n_bottom), dtype=np.float32)
    for b in range(B):
        for t in rng.choice(np.arange(2, hier.n_top), size=rng.choice([1, 2, 3], p=[0.69, 0.30, 0.01]), replace=False):
            ids_t = hier.top2bottom[int(t)]
            cand = [x for x in ids_t if x not in hier.none_bottoms]
            labels[b, rng.choice(cand)] = 1
        if rng.rand() < 0.1:
            labels[b, 1] = 1                                                 # <unk> label column is a live BCE term
    batch["labels"] = torch.from_numpy(labels)
    return batch


def run_reference(nb, make_model, BertAdam, reverse_top2bottom, cfg, hier, params, batch, hp, n_steps):
    enc = hf_encoder(cfg)
    opt = Namespace(pretrained_model=enc, dropout=0.0, device=torch.device("cpu"), score_util="none", sent_repr="cls",
                    cls_type="stc", top2bottom_dict=hier.top2bottom, label_vocab_size=hier.n_bottom,
                    pre_trained_model="xlm-roberta" if cfg.kind == "xlm-roberta" else "bert",
                    add_l2_loss=hp["add_l2_loss"], class_loss_function=torch.nn.BCELoss(reduction="sum"),
                    ce_loss_function=torch.nn.NLLLoss(reduction="sum"), mse_loss_function=torch.nn.MSELoss())
    model = make_model(opt)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k or "token_type_ids" in k for k in missing), missing
    model.eval()                                                             # HF-internal dropout off; head p = 0
    memory = dict(top2bottom_dict=hier.top2bottom, bottom2top_mat=reverse_top2bottom(hier.top2bottom))
    groups = []
    for n, p in model.named_parameters():                                    # n_best_asr_bert.py:540-550
        lr_p, wd = O.param_hyper(n, hp["lr"], hp["bert_lr"])
        groups.append(dict(params=p, weight_decay=wd, lr=lr_p))
    optim = BertAdam(groups, lr=hp["lr"], warmup=hp["warmup"], t_total=hp["t_total"])
    outs = []
    import contextlib
    import io
    for _ in range(n_steps):
        optim.zero_grad()
        top, bottoms, final, asr, trans = model(opt, batch["ids"], batch.get("trans_ids"), seg_ids=batch.get("seg"),
                                                trans_seg_ids=batch.get("trans_seg"), classifier_input_type="asr")
        with contextlib.redirect_stdout(io.StringIO()):                      # n_best_asr_bert.py:169 prints
            rec, total = nb.cal_total_loss(top, bottoms, final, batch["labels"], memory, opt, asr, trans)
        total.backward()
        grads = {n: (p.grad.clone() if p.grad is not None else None) for n, p in model.named_parameters()}
        outs.append(dict(top=top.detach().clone(), bottoms={k: v.detach().clone() for k, v in bottoms.items()},
                         final=final.detach().clone(), asr=asr.detach().clone(),
                         trans=None if trans is None else trans.detach().clone(), total=float(total), rec=rec, grads=grads))
        optim.step()
    post = {n: p.detach().clone() for n, p in model.named_parameters()}
    return outs, post


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def case(name, cfg, hier, refmods, B, max_len, seed, add_l2, n_steps=3, keep_grads=()):
    nb, make_model, BertAdam, reverse_top2bottom, _ = refmods
    params = O.init_params(cfg, hier, seed=seed, style="perturbed")
    batch = synth_batch(cfg, hier, B, max_len, seed + 1)
    hp = dict(lr=3e-4, bert_lr=1e-4, warmup=0.1, t_total=8, add_l2_loss=add_l2)
    ref_outs, ref_post = run_reference(nb, make_model, BertAdam, reverse_top2bottom, cfg, hier, params, batch, hp, n_steps)

    # oracle on the same weights / inputs
    oparams = {k: v.clone() for k, v in params.items()}
    state = {}
    stats = {}
    for s in range(n_steps):
        terms, grads, (top, bottoms, final, asr, trans) = O.train_step(oparams, cfg, hier, batch, state, hp)
        r = ref_outs[s]
        stats.setdefault("top", []).append(rel(top.detach(), r["top"]))
        stats.setdefault("final", []).append(rel(final.detach(), r["final"]))
        stats.setdefault("asr", []).append(rel(asr.detach(), r["asr"]))
        stats.setdefault("loss", []).append(abs(terms["total"] - r["total"]) / abs(r["total"]))
        worst = 0.0
        for n, g in r["grads"].items():
            if g is None:
                assert grads[n] is None or float(grads[n].abs().max()) == 0.0, n
                continue
            if "attention.self.key.bias" in n:                               # analytically zero
                continue
            worst = max(worst, rel(grads[n], g))
        stats.setdefault("grad", []).append(worst)
    post_err = max(float((oparams[n] - ref_post[n]).abs().max()) for n in ref_post if "attention.self.key.bias" not in n)
    print("%-14s oracle-vs-reference: top %.1e final %.1e cls %.1e loss %.1e grad %.1e  post-step |dp| %.1e" % (
        name, max(stats["top"]), max(stats["final"]), max(stats["asr"]), max(stats["loss"]), max(stats["grad"]), post_err))
    assert max(stats["top"]) < 1e-4 and max(stats["final"]) < 1e-4 and max(stats["asr"]) < 1e-4
    assert max(stats["loss"]) < 1e-5 and max(stats["grad"]) < 2e-3 and post_err < 2e-5

    # decode parity (pred_one_sample)
    r0 = ref_outs[0]
    memory = dict(top2bottom_dict=hier.top2bottom, idx2label={i: ("x-NONE" if i in hier.none_bottoms else "l%d" % i)
                                                               for i in range(hier.n_bottom)})
    dec_ref = np.zeros((B, hier.n_bottom), dtype=np.uint8)
    for i, ts in enumerate(r0["top"].tolist()):
        for lbl in nb.pred_one_sample(i, ts, r0["bottoms"], memory, None):
            dec_ref[i, int(lbl[1:])] = 1
    dec_or = O.decode(hier, r0["top"], r0["bottoms"])
    assert (dec_ref == dec_or).all()

    fx = dict(cfg=json.dumps(cfg.__dict__), seed=seed, B=B, hp=json.dumps(hp), n_steps=n_steps,
              ids=batch["ids"].numpy(), seg=batch["seg"].numpy(), trans_ids=batch["trans_ids"].numpy(),
              trans_seg=batch["trans_seg"].numpy(), labels=batch["labels"].numpy(),
              weight_checksum=np.array([float(sum(v.double().abs().sum() for v in params.values()))]),
              decode=dec_ref)
    for s, r in enumerate(ref_outs):
        fx["top_%d" % s] = r["top"].numpy()
        fx["final_%d" % s] = r["final"].numpy()
        fx["bottom_%d" % s] = torch.cat([r["bottoms"]["lin_%d" % k] for k in hier.group_tops], 1).numpy()
        fx["asr_%d" % s] = r["asr"].numpy()
        fx["trans_%d" % s] = r["trans"].numpy()
        fx["total_%d" % s] = np.array([r["total"]])
        fx["rec_%d" % s] = np.array([r["rec"]])
        fx["gradnorm_%d" % s] = np.array([0.0 if g is None else float(g.double().norm()) for g in r["grads"].values()])
        for n in keep_grads:
            fx["grad_%d_%s" % (s, n)] = r["grads"][n].numpy()
    fx["grad_names"] = np.array(list(ref_outs[0]["grads"].keys()))
    fx["post_norm"] = np.array([float(v.double().norm()) for v in ref_post.values()])
    for n in keep_grads:
        fx["post_%s" % n] = ref_post[n].numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **fx)


def packing_case(prepare_inputs_for_roberta, m):
    """A2 golden: the reference's own prepare_inputs_for_roberta on real lines of the shipped `valid` file with a
    deterministic fake tokenizer (no vocabularies exist offline) -> padded tensors the packed layout must invert."""

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_tokenizer import FakeTok, FakeXlmrTok

    lines = open(os.path.join(REF, "dstc2_data/processed_data/raw/valid")).read().strip().split("\n")[:24]
    raw_in = [l.split("\t<=>\t")[0].strip().split(" ") for l in lines]
    out = {}
    cases = (("default", FakeTok(), dict(without_system_act=False, tod_pre_trained_model=None, pre_trained_model="bert")),
             ("nosys", FakeTok(), dict(without_system_act=True, tod_pre_trained_model=None, pre_trained_model="bert")),
             ("tod", FakeTok(), dict(without_system_act=False, tod_pre_trained_model="tod-bert", pre_trained_model=None)),
             ("xlmr", FakeXlmrTok(), dict(without_system_act=False, tod_pre_trained_model=None, pre_trained_model="xlm-roberta")))
    for tag, tok, kw in cases:
        opt = Namespace(**kw)
        ids, seg, lens = prepare_inputs_for_roberta(raw_in, tok, opt, torch.device("cpu"))
        out["ids_" + tag] = ids.numpy()
        out["lens_" + tag] = np.array(lens)
        if seg is not None:
            out["seg_" + tag] = seg.numpy()
    out["raw_in"] = np.array([" ".join(x) for x in raw_in])
    np.savez_compressed(os.path.join(GOLD, "packing_valid24.npz"), **out)
    print("packing fixture: ids", out["ids_default"].shape, "lens", out["lens_default"][:8])


def epoch_case(nb, m):
    """A1 / A11 / A10 / A12 golden: on 48 real lines of the shipped `valid` file run the reference's own collate_fn
    (utils/dataset/tod_asr_util.py:86-132), and on seeded random scores its pred_one_sample, filter_informative and
    update_f1 / compute_f1 -> gold multi-hots, predicted label strings and the epoch counters our device metrics must hit."""
    from utils.dataset.tod_asr_util import collate_fn
    from utils.fscore import update_f1, compute_f1
    lines = open(os.path.join(REF, "dstc2_data/processed_data/raw/valid")).read().strip("\n").split("\n")
    pick = [i for i, l in enumerate(lines) if l.split("\t<=>\t")[2].strip()][:40] + \
        [i for i, l in enumerate(lines) if any(x not in m["label2idx"] for x in l.split("\t<=>\t")[2].strip().split(";"))][:8]
    batch = []
    for i in pick:
        a, t, l = lines[i].strip("\n\r").split("\t<=>\t")
        batch.append((a.strip().split(" "), t.strip().split(" "), [] if len(l) == 0 else l.strip().split(";")))
    labels, in_seqs, trans_seqs, label_lists = collate_fn(batch, m, 512, torch.device("cpu"))
    hier = O.Hierarchy({int(k): [int(x) for x in v] for k, v in m["top2bottom_dict"].items()},
                       [int(i) for i, l in m["idx2label"].items() if l.endswith("NONE")])
    g = torch.Generator().manual_seed(77)
    B = len(batch)
    top = torch.rand(B, hier.n_top, generator=g) * 0.51          # ~2 % spurious act-slots above the 0.5 threshold
    # make the gold act-slots likely to fire so that TP is well populated
    b2t = m["bottom2top_mat"] if "bottom2top_mat" in m else None
    bottoms, flat = {}, []
    for k in sorted(hier.top2bottom):
        ids = hier.top2bottom[k]
        if len(ids) > 1:
            bottoms["lin_%d" % k] = torch.softmax(3 * torch.randn(B, len(ids), generator=g), -1)
            flat.append(bottoms["lin_%d" % k])
    for i, ll in enumerate(label_lists):           # bias towards the gold labels (two thirds of them)
        for l in ll:
            if l in m["label2idx"] and (i + len(l)) % 3 != 0:
                bi = m["label2idx"][l]
                for k, ids in hier.top2bottom.items():
                    if bi in ids:
                        top[i, k] = 0.9
                        if len(ids) > 1:
                            bottoms["lin_%d" % k][i] = 0.01
                            bottoms["lin_%d" % k][i, ids.index(bi)] = 0.9
    opt = Namespace()
    ontology = json.load(open(os.path.join(REF, "dstc2_data/processed_data/ontology_dstc2.json"))) \
        if os.path.exists(os.path.join(REF, "dstc2_data/processed_data/ontology_dstc2.json")) else \
        dict(informable=dict(food=["a", "b"], area=["x", "y"], pricerange=["p", "q"], name=["n"]))
    preds, preds_f, counts, counts_f = [], [], [0, 0, 0, 0], [0, 0, 0, 0]
    for i, (ts, gold) in enumerate(zip(top.tolist(), label_lists)):
        pc = nb.pred_one_sample(i, ts, bottoms, m, opt)
        preds.append(pc)
        counts[:3] = update_f1(pc, gold, *counts[:3])
        counts[3] += int(set(pc) == set(gold))
        pf, gf = nb.filter_informative(pc, ontology), nb.filter_informative(gold, ontology)
        preds_f.append(pf)
        counts_f[:3] = update_f1(pf, gf, *counts_f[:3])
        counts_f[3] += int(set(pf) == set(gold if False else gf))
    out = dict(labels=labels.numpy(), top=top.numpy(), bottom=torch.cat(flat, 1).numpy(),
               counts=np.array(counts), counts_filtered=np.array(counts_f),
               prf=np.array(compute_f1(*counts[:3])), prf_filtered=np.array(compute_f1(*counts_f[:3])),
               meta=json.dumps(dict(label_lists=label_lists, preds=preds, preds_filtered=preds_f, ontology=ontology,
                                    label2idx=m["label2idx"], raw_in=[" ".join(x) for x in in_seqs],
                                    raw_trans=[" ".join(x) for x in trans_seqs])))
    np.savez_compressed(os.path.join(GOLD, "epoch_valid48.npz"), **out)
    print("epoch fixture: B", B, "counts", counts, "filtered", counts_f, "unk golds", int(labels[:, 1].sum()))


def export_coverage():
    """`--coverage` stratified sampler (utils/dataset/tod_asr_util.py:12-39, pandas) run by the live reference on the
    shipped valid file -> tests/golden/coverage_valid.npz: the integer-coded label list of every utterance (all the
    sampler looks at) and, per coverage value, which utterances the reference kept, in its order."""
    import contextlib
    import io
    import_reference()
    import utils.dataset.tod_asr_util as tod
    fn = os.path.join(REF, "dstc2_data/processed_data/raw/valid")
    asr, trans, labels = tod.read_wcn_data(fn)
    codes, code_of = [], {}
    for l in labels:
        codes.append(code_of.setdefault(tuple(l), len(code_of)))
    pos = {}
    for i, (a, t) in enumerate(zip(asr, trans)):
        pos.setdefault((tuple(a), tuple(t), tuple(labels[i])), []).append(i)
    out = dict(label_code=np.asarray(codes, dtype=np.int32))
    for cov in (0.1, 0.25, 0.5):
        with contextlib.redirect_stdout(io.StringIO()):
            a2, t2, l2 = tod.read_wcn_data(fn, cov)
        used, idx = {}, []
        for a, t, l in zip(a2, t2, l2):            # map the sampled rows back to line numbers (duplicates in file order)
            key = (tuple(a), tuple(t), tuple(l))
            k = used.get(key, 0)
            idx.append(pos[key][k] if k < len(pos[key]) else pos[key][-1])
            used[key] = k + 1
        out["count_%g" % cov] = np.asarray([len(a2)], dtype=np.int64)
        out["labels_%g" % cov] = np.asarray([code_of[tuple(l)] for l in l2], dtype=np.int32)
        out["index_%g" % cov] = np.asarray(idx, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "coverage_valid.npz"), **out)
    print("coverage fixture:", {k: v.shape for k, v in out.items()})


def main():
    if "--coverage-only" in sys.argv:
        return export_coverage()
    if "--epoch-only" in sys.argv:
        refmods = import_reference()
        m = torch.load(os.path.join(REF, "dstc2_data/processed_data/raw/memory.pt"))
        return epoch_case(refmods[0], m)
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    refmods = import_reference()
    hier, m = export_hierarchy()
    keep = ("clf.top_linear_layer.weight", "clf.linear_layers.lin_2.bias",
            "bert_encoder.encoder.layer.0.attention.self.query.bias",
            "bert_encoder.encoder.layer.1.output.LayerNorm.weight",
            "bert_encoder.embeddings.token_type_embeddings.weight")
    small = dict(vocab_size=1200, layers=2, max_position=96)
    case("bert_l2_small", O.EncoderConfig.bert_base(**small), hier, refmods, B=6, max_len=40, seed=11, add_l2=True,
         keep_grads=keep)
    case("bert_nol2_small", O.EncoderConfig.bert_base(**small), hier, refmods, B=5, max_len=33, seed=12, add_l2=False,
         keep_grads=keep)
    case("xlmr_l2_small", O.EncoderConfig.xlmr_base(vocab_size=1200, layers=2, max_position=98), hier, refmods, B=4,
         max_len=30, seed=13, add_l2=True, keep_grads=keep[:4])
    packing_case(refmods[4], m)
    epoch_case(refmods[0], m)


if __name__ == "__main__":
    main()
