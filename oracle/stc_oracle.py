"""CPU oracle for the N-Best-ASR-Transformer hot path — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32 restatement of the reference algorithm for the path BASELINE.json names. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the product
package (n-best-asr-transformer_b200/) never does and has no CPU fallback.

Parity pin: the reference has no tests or golden vectors of its own, so this oracle is pinned
against the reference ITSELF run in the build container: oracle/make_golden.py imports the unmodified reference
modules from /root/reference together with the installed HuggingFace encoder, feeds both the same weights and
inputs, asserts agreement (fp32 round-off) and stores the REFERENCE's outputs under tests/golden/. tests/
test_oracle_golden.py re-checks this oracle against those stored vectors wherever the tests run.

Each function cites the reference lines it follows (paths relative to /root/reference; `hf:` = the third-party
transformers package the reference calls, pinned 2.3.0 in requirements.txt:5, 5.5.0 installed).
"""
import math
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class EncoderConfig:
    """bert-base-uncased / xlm-roberta-base hyper-parameters (hf: BertConfig / XLMRobertaConfig defaults)."""
    kind: str = "bert"                 # "bert" | "xlm-roberta"   (n_best_asr_bert.py:33-37)
    vocab_size: int = 30522
    hidden: int = 768                  # models/model.py:30 hard-codes fea_dim = 768
    layers: int = 12
    heads: int = 12
    intermediate: int = 3072
    max_position: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12
    pad_token_id: int = 0              # nn.Embedding padding_idx of the word table

    @staticmethod
    def bert_base(**kw):
        return EncoderConfig(**kw)

    @staticmethod
    def xlmr_base(**kw):
        d = dict(kind="xlm-roberta", vocab_size=250002, max_position=514, type_vocab=1, ln_eps=1e-5, pad_token_id=1)
        d.update(kw)
        return EncoderConfig(**d)


class Hierarchy:
    """memory['top2bottom_dict'] (n_best_asr_bert.py:489-496) in the flattened form the fused head uses.

    Columns of the fused logit matrix: n_top act-slot columns, then for every multi-way top id (ascending) its
    n_k value columns (hierarchical_classifier.py:15-24: one Linear per top label with >= 2 bottoms).
    """

    def __init__(self, top2bottom, none_bottoms=()):
        self.top2bottom = {int(k): [int(x) for x in v] for k, v in top2bottom.items()}
        self.n_top = len(self.top2bottom)
        self.n_bottom = sum(len(v) for v in self.top2bottom.values())
        self.group_tops = [k for k in sorted(self.top2bottom) if len(self.top2bottom[k]) >= 2]
        self.n_groups = len(self.group_tops)
        self.grp_off = [self.n_top]
        for k in self.group_tops:
            self.grp_off.append(self.grp_off[-1] + len(self.top2bottom[k]))
        self.n_cols = self.grp_off[-1]
        self.col_group = [0] * self.n_top
        self.col_bottom = [self.top2bottom[i][0] if len(self.top2bottom[i]) == 1 else -1 for i in range(self.n_top)]
        for g, k in enumerate(self.group_tops, start=1):
            self.col_group += [g] * len(self.top2bottom[k])
            self.col_bottom += list(self.top2bottom[k])
        self.none_bottoms = set(int(x) for x in none_bottoms)     # bottom ids whose label ends with 'NONE'
        self.none_col = [1 if (c >= self.n_top and self.col_bottom[c] in self.none_bottoms) else 0
                         for c in range(self.n_cols)]
        b2t = {}
        for k, vs in self.top2bottom.items():
            for v in vs:
                b2t[v] = k
        self.b2t = [b2t[i] for i in range(self.n_bottom)]

    def b2t_matrix(self):
        """utils/STC_util.py:10-26 reverse_top2bottom."""
        m = torch.zeros(self.n_bottom, self.n_top)
        m[torch.arange(self.n_bottom), torch.tensor(self.b2t)] = 1
        return m

    @staticmethod
    def from_json(d):
        return Hierarchy({int(k): v for k, v in d["top2bottom"].items()}, d.get("none_bottoms", ()))


# ----------------------------------------------------------------------------------------------------------------
# parameters (state_dict names of models/model.py:19,32 wrapping the HF encoder)
# ----------------------------------------------------------------------------------------------------------------
def param_shapes(cfg, hier):
    H, I = cfg.hidden, cfg.intermediate
    s = OrderedDict()
    e = "bert_encoder.embeddings."
    s[e + "word_embeddings.weight"] = (cfg.vocab_size, H)
    s[e + "position_embeddings.weight"] = (cfg.max_position, H)
    s[e + "token_type_embeddings.weight"] = (cfg.type_vocab, H)
    s[e + "LayerNorm.weight"] = (H,)
    s[e + "LayerNorm.bias"] = (H,)
    for l in range(cfg.layers):
        p = "bert_encoder.encoder.layer.%d." % l
        for n in ("query", "key", "value"):
            s[p + "attention.self.%s.weight" % n] = (H, H)
            s[p + "attention.self.%s.bias" % n] = (H,)
        s[p + "attention.output.dense.weight"] = (H, H)
        s[p + "attention.output.dense.bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,)
        s[p + "attention.output.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (I, H)
        s[p + "intermediate.dense.bias"] = (I,)
        s[p + "output.dense.weight"] = (H, I)
        s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,)
        s[p + "output.LayerNorm.bias"] = (H,)
    s["bert_encoder.pooler.dense.weight"] = (H, H)
    s["bert_encoder.pooler.dense.bias"] = (H,)
    s["clf.top_linear_layer.weight"] = (hier.n_top, H)
    s["clf.top_linear_layer.bias"] = (hier.n_top,)
    for k in hier.group_tops:
        n = len(hier.top2bottom[k])
        s["clf.linear_layers.lin_%d.weight" % k] = (n, H)
        s["clf.linear_layers.lin_%d.bias" % k] = (n,)
    return s


def init_params(cfg, hier, seed=999, style="hf"):
    """Deterministic random-init weights (no checkpoints exist offline).

    style="hf": N(0, 0.02) matrices/tables, zero biases, LN = (1, 0), padding rows zeroed — the HF initialiser's
    distribution; style="perturbed": additionally random biases and LN parameters so that every term of the math
    is exercised by parity tests.
    """
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape in param_shapes(cfg, hier).items():
        if name.endswith("LayerNorm.weight"):
            t = torch.ones(shape)
            if style == "perturbed":
                t = t + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = torch.zeros(shape)
            if style == "perturbed":
                t = 0.05 * torch.randn(shape, generator=g)
        else:
            std = 0.05 if (style == "perturbed" and name.startswith("clf.")) else 0.02
            t = std * torch.randn(shape, generator=g)
        out[name] = t
    out["bert_encoder.embeddings.word_embeddings.weight"][cfg.pad_token_id].zero_()
    if cfg.kind == "xlm-roberta":
        out["bert_encoder.embeddings.position_embeddings.weight"][1].zero_()
    return out


# ----------------------------------------------------------------------------------------------------------------
# A2: packed layout (restates what un-padding the tensors of utils/bert_xlnet_inputs.py:91-102 must give)
# ----------------------------------------------------------------------------------------------------------------
def pack_batch(ids, seg_ids, kind):
    """ids [B,S] int64 right-padded, seg_ids [B,S] or None -> dict of numpy arrays (bit-exact contract).

    length_i = 1 + last index with id > 0 (the reference's attention mask is `input_ids > 0`, models/model.py:43):
    BERT rows lose their pad=0 tail; XLM-R rows keep their <pad>=1 tail because the reference leaves those
    positions attendable. key_valid = ids > 0. Position ids: BERT arange; XLM-R
    cumsum(ids != 1) * (ids != 1) + 1 (hf: modeling_xlm_roberta.py:147-160).
    """
    ids = np.asarray(ids, dtype=np.int64)
    B, S = ids.shape
    valid = ids > 0
    lens = np.zeros(B, dtype=np.int32)
    for b in range(B):
        nz = np.nonzero(valid[b])[0]
        lens[b] = (nz[-1] + 1) if nz.size else 0
    cu = np.zeros(B + 1, dtype=np.int32)
    cu[1:] = np.cumsum(lens)
    T = int(cu[-1])
    tokens = np.zeros(T, dtype=np.int32)
    seg = np.zeros(T, dtype=np.uint8)
    pos = np.zeros(T, dtype=np.int32)
    seq_of = np.zeros(T, dtype=np.int32)
    key_valid = np.zeros(T, dtype=np.uint8)
    for b in range(B):
        L = lens[b]
        sl = slice(cu[b], cu[b] + L)
        tokens[sl] = ids[b, :L]
        if seg_ids is not None:
            seg[sl] = np.asarray(seg_ids)[b, :L]
        if kind == "xlm-roberta":
            nonpad = (ids[b] != 1).astype(np.int64)
            pos[sl] = (np.cumsum(nonpad) * nonpad + 1)[:L]
        else:
            pos[sl] = np.arange(L)
        seq_of[sl] = b
        key_valid[sl] = valid[b, :L]
    return dict(lens=lens, cu_seqlens=cu, tokens=tokens, seg=seg, pos=pos, seq_of=seq_of, key_valid=key_valid, T=T)


# ----------------------------------------------------------------------------------------------------------------
# A4: encoder (hf: modeling_bert.py:53-112 embeddings, :143-207 self-attention, :287-298, :330-356, :424-453)
# ----------------------------------------------------------------------------------------------------------------
def _ln(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def encoder_forward(params, cfg, input_ids, token_type_ids=None, drop=None, prefix="bert_encoder."):
    """input_ids [B,S] int64 -> last hidden state [B,S,H]; attention mask = input_ids > 0 (models/model.py:43-45).

    `drop` is None (dropout off: model.eval() or p = 0) or a callable (tensor, site) -> tensor used by the CPU
    baseline timing to apply torch dropout where HF does (embeddings, attention probs, both dense outputs).
    """
    P = lambda n: params[prefix + n]
    B, S = input_ids.shape
    H, nh = cfg.hidden, cfg.heads
    dh = H // nh
    key_ok = input_ids > 0                                                   # [B,S] bool, key mask only
    if cfg.kind == "xlm-roberta":
        nonpad = (input_ids != 1).long()
        pos_ids = torch.cumsum(nonpad, dim=1) * nonpad + 1                   # hf xlm_roberta :147-160
        tt = torch.zeros_like(input_ids)                                     # models/model.py:42-43: no token types
    else:
        pos_ids = torch.arange(S).unsqueeze(0).expand(B, S)
        tt = token_type_ids if token_type_ids is not None else torch.zeros_like(input_ids)
    # nn.Embedding(padding_idx=...) rows never receive gradient: word[pad] always, position[1] on XLM-R
    x = (F.embedding(input_ids, P("embeddings.word_embeddings.weight"), padding_idx=cfg.pad_token_id)
         + F.embedding(tt, P("embeddings.token_type_embeddings.weight"))
         + F.embedding(pos_ids, P("embeddings.position_embeddings.weight"),
                       padding_idx=1 if cfg.kind == "xlm-roberta" else None))
    x = _ln(x, P("embeddings.LayerNorm.weight"), P("embeddings.LayerNorm.bias"), cfg.ln_eps)
    if drop is not None:
        x = drop(x, "emb")
    neg = torch.zeros(B, 1, 1, S).masked_fill(~key_ok[:, None, None, :], float("-inf"))
    for l in range(cfg.layers):
        lp = "encoder.layer.%d." % l
        def lin(t, n):
            return t @ P(lp + n + ".weight").t() + P(lp + n + ".bias")
        q = lin(x, "attention.self.query").view(B, S, nh, dh).transpose(1, 2)
        k = lin(x, "attention.self.key").view(B, S, nh, dh).transpose(1, 2)
        v = lin(x, "attention.self.value").view(B, S, nh, dh).transpose(1, 2)
        scores = q @ k.transpose(-1, -2) / math.sqrt(dh) + neg
        probs = torch.softmax(scores, dim=-1)
        if drop is not None:
            probs = drop(probs, "attn")
        ctx = (probs @ v).transpose(1, 2).reshape(B, S, H)
        a = lin(ctx, "attention.output.dense")
        if drop is not None:
            a = drop(a, "hidden")
        x = _ln(a + x, P(lp + "attention.output.LayerNorm.weight"), P(lp + "attention.output.LayerNorm.bias"), cfg.ln_eps)
        h = gelu_erf(lin(x, "intermediate.dense"))
        o = lin(h, "output.dense")
        if drop is not None:
            o = drop(o, "hidden")
        x = _ln(o + x, P(lp + "output.LayerNorm.weight"), P(lp + "output.LayerNorm.bias"), cfg.ln_eps)
    return x


# ----------------------------------------------------------------------------------------------------------------
# A5: hierarchical STC head (models/modules/hierarchical_classifier.py:35-60)
# ----------------------------------------------------------------------------------------------------------------
def head_forward(params, hier, f, drop=None):
    """f [B,768] -> (top_scores [B,n_top], {lin_k: [B,n_k]}, final_scores [B,n_bottom])."""
    fd = (lambda: f) if drop is None else (lambda: drop(f, "head"))          # fresh mask per Linear call (:41,:46)
    top = torch.sigmoid(fd() @ params["clf.top_linear_layer.weight"].t() + params["clf.top_linear_layer.bias"])
    bottoms = OrderedDict()
    for k in hier.group_tops:
        z = fd() @ params["clf.linear_layers.lin_%d.weight" % k].t() + params["clf.linear_layers.lin_%d.bias" % k]
        bottoms["lin_%d" % k] = torch.softmax(z, dim=1)
    cols = [None] * hier.n_bottom
    for i in range(hier.n_top):
        ids = hier.top2bottom[i]
        if len(ids) >= 2:
            prod = top[:, i:i + 1] * bottoms["lin_%d" % i]
            for j, b in enumerate(ids):
                cols[b] = prod[:, j]
        else:
            cols[ids[0]] = top[:, i]
    final = torch.stack(cols, dim=1)
    return top, bottoms, final


def model_forward(params, cfg, hier, input_ids, trans_input_ids=None, seg_ids=None, trans_seg_ids=None, drop=None,
                  classifier_input_type="asr"):
    """models/model.py:35-73: both streams through the shared encoder, CLS row, head on the ASR stream."""
    asr = encoder_forward(params, cfg, input_ids, seg_ids, drop)[:, 0, :]
    trans = None
    if trans_input_ids is not None:
        trans = encoder_forward(params, cfg, trans_input_ids, trans_seg_ids, drop)[:, 0, :]
    lin_in = trans if classifier_input_type == "transcript" else asr
    top, bottoms, final = head_forward(params, hier, lin_in, drop)
    return top, bottoms, final, asr, trans


# ----------------------------------------------------------------------------------------------------------------
# A6: losses (n_best_asr_bert.py:145-195, utils/STC_util.py:4-51)
# ----------------------------------------------------------------------------------------------------------------
def _bce_sum(p, t):
    """torch.nn.BCELoss(reduction='sum') (n_best_asr_bert.py:572): forward -(t*max(log p,-100) + (1-t)*max(log(1-p),-100)),
    backward the ATen closed form (p-t)/max(p(1-p),1e-12). The torch op itself is used so that saturated scores
    differentiate exactly as they do for the reference (a hand-written log+clamp gives 0/0 there)."""
    return F.binary_cross_entropy(p, t, reduction="sum")


def total_loss(hier, top, bottoms, final, labels, asr_cls=None, trans_cls=None, add_l2_loss=False):
    """Returns (total, dict of the four terms). All class terms are sums over the batch."""
    terms = OrderedDict()
    total = 0.0
    if add_l2_loss and asr_cls is not None and trans_cls is not None:
        terms["mse"] = ((asr_cls - trans_cls) ** 2).mean()                   # nn.MSELoss() default mean (:574)
        total = total + terms["mse"]
    terms["bce_final"] = _bce_sum(final, labels)
    total = total + terms["bce_final"]
    top_labels = labels @ hier.b2t_matrix()                                  # STC_util.py:4-7
    terms["bce_top"] = _bce_sum(top, top_labels)
    total = total + terms["bce_top"]
    ces = []
    for k in hier.group_tops:
        ids = hier.top2bottom[k]
        sub = labels[:, ids]
        assert bool((sub.sum(1) <= 1).all())                                 # STC_util.py:34
        tgt = sub.argmax(1)
        tgt = torch.where(sub.sum(1) == 0, torch.full_like(tgt, len(ids) - 1), tgt)   # STC_util.py:36-49
        logp = torch.log(bottoms["lin_%d" % k] + 1e-12)                      # n_best_asr_bert.py:154
        ces.append(-logp[torch.arange(logp.shape[0]), tgt].sum())            # NLLLoss(sum)
    terms["ce"] = sum(ces) / len(ces)
    total = total + terms["ce"]
    return total, terms


def decode(hier, top, bottoms):
    """pred_one_sample (n_best_asr_bert.py:198-215) for a whole batch -> uint8 bitmap [B,n_bottom]."""
    B = top.shape[0]
    out = np.zeros((B, hier.n_bottom), dtype=np.uint8)
    topn = top.detach().numpy()
    for b in range(B):
        for ti in range(hier.n_top):
            if topn[b, ti] > 0.5:
                ids = hier.top2bottom[ti]
                if len(ids) == 1:
                    out[b, ids[0]] = 1
                else:
                    j = int(bottoms["lin_%d" % ti][b].detach().numpy().argmax())
                    if ids[j] not in hier.none_bottoms:
                        out[b, ids[j]] = 1
    return out


# ----------------------------------------------------------------------------------------------------------------
# A8/A9: parameter groups + BertAdam (n_best_asr_bert.py:535-550, models/optimization.py:162-171,237-302)
# ----------------------------------------------------------------------------------------------------------------
def param_hyper(name, lr, bert_lr):
    no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
    wd = 0.0 if any(nd in name for nd in no_decay) else 0.01
    return (bert_lr if "bert_encoder" in name else lr), wd


def warmup_linear(step, t_total, warmup):
    if t_total < 0:
        return 1.0
    x = float(step) / float(t_total)
    if x < warmup:
        return x / warmup
    return max((x - 1.0) / (warmup - 1.0), 0.0)


def bertadam_step(params, grads, state, lr, bert_lr, warmup, t_total, b1=0.9, b2=0.999, eps=1e-6, max_grad_norm=1.0):
    """In-place update of `params` (dict name -> fp32 tensor); tensors whose grad is None are skipped entirely."""
    for name, p in params.items():
        g = grads.get(name)
        if g is None:
            continue
        st = state.setdefault(name, dict(step=0, m=torch.zeros_like(p), v=torch.zeros_like(p)))
        lr_p, wd = param_hyper(name, lr, bert_lr)
        if max_grad_norm > 0:                                               # clip_grad_norm_ on ONE tensor (:270-271)
            n = float(torch.linalg.vector_norm(g.double()).float())
            coef = max_grad_norm / (n + 1e-6)
            if coef < 1.0:
                g = g * coef
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        upd = st["m"] / (st["v"].sqrt() + eps)
        if wd > 0:
            upd = upd + wd * p
        p.add_(upd, alpha=-(lr_p * warmup_linear(st["step"], t_total, warmup)))
        st["step"] += 1


# ----------------------------------------------------------------------------------------------------------------
# whole training step, used as the checker and as the CPU baseline ("port")
# ----------------------------------------------------------------------------------------------------------------
def train_step(params, cfg, hier, batch, opt_state, hp, drop=None):
    """batch: dict(ids, seg, trans_ids, trans_seg, labels). Returns (loss terms dict, grads dict)."""
    leaves = {k: v.detach().requires_grad_(True) for k, v in params.items()}
    top, bottoms, final, asr, trans = model_forward(leaves, cfg, hier, batch["ids"], batch.get("trans_ids"),
                                                    batch.get("seg"), batch.get("trans_seg"), drop)
    total, terms = total_loss(hier, top, bottoms, final, batch["labels"], asr, trans, hp.get("add_l2_loss", False))
    total.backward()
    grads = {k: v.grad for k, v in leaves.items()}
    if opt_state is not None:
        with torch.no_grad():
            bertadam_step(params, grads, opt_state, hp["lr"], hp["bert_lr"], hp["warmup"], hp["t_total"])
    return (dict(total=float(total.detach()), **{k: float(v.detach()) for k, v in terms.items()}), grads,
            (top, bottoms, final, asr, trans))


# ----------------------------------------------------------------------------------------------- dropout mask restatement
# The reference draws its dropout masks from torch's Philox stream (nn.Dropout inside HF BERT, hierarchical_classifier.py:
# 41,46); any independent Bernoulli(1-p) mask is an equally valid realisation, so the CUDA path uses its own counter-based
# hash (csrc/ptx.cuh dropout_quad) that forward and backward can both regenerate. This numpy restatement pins it bit-exactly.
def mix_seed(s):
    """ops._seed: murmur3 finaliser applied on the host to the structured (step, layer, site) seed."""
    s = int(s) & 0xFFFFFFFF
    s ^= s >> 16
    s = (s * 0x85EBCA6B) & 0xFFFFFFFF
    s ^= s >> 13
    s = (s * 0xC2B2AE35) & 0xFFFFFFFF
    s ^= s >> 16
    return s


def dropout_lanes(seed, quad_idx):
    """csrc/ptx.cuh dropout_quad: four 16-bit lanes per 32-bit counter. `seed` is the MIXED seed; returns int64 [..., 4]."""
    import numpy as np
    m32 = np.uint64(0xFFFFFFFF)
    fold = lambda m: ((m & m32) ^ (m >> np.uint64(32))) & m32
    q = np.asarray(quad_idx).astype(np.uint64) & m32
    x = (fold(q * np.uint64(0x9E3779B1)) ^ np.uint64(seed)) & m32
    x ^= x >> np.uint64(16)
    a, b = fold(x * np.uint64(0x85EBCA6B)), fold(x * np.uint64(0xC2B2AE35))
    s16, m16 = np.uint64(16), np.uint64(0xFFFF)
    return np.stack([a & m16, a >> s16, b & m16, b >> s16], -1).astype(np.int64)


def dropout_threshold(p):
    t = p * 65536.0 + 0.5
    return 0 if p <= 0 else (65535 if t >= 65535.0 else int(t))


def dropout_keep_mask(seed, n, p):
    """Keep mask of elements 0 .. n-1 of a row-major tensor (element e = lane e % 4 of quad e // 4); seed un-mixed."""
    import numpy as np
    lanes = dropout_lanes(mix_seed(seed), np.arange((n + 3) // 4)).reshape(-1)[:n]
    return lanes >= dropout_threshold(p)


def attn_dropout_keep_mask(seed, heads, T, tq, L, p):
    """Keep mask [heads, L] of the attention probabilities of global query token tq over the L keys of its sequence
    (csrc/ptx.cuh attn_quad / attn_lane)."""
    import numpy as np
    j = np.arange(L)
    h = np.arange(heads)[:, None]
    quad = ((h * T + tq) * 128 + ((j >> 4) << 2) + ((j & 7) >> 1))[..., None]
    lane = ((j & 1) | ((j >> 2) & 2))[None, :, None]
    lanes = dropout_lanes(mix_seed(seed), quad[..., 0])
    return np.take_along_axis(lanes, np.broadcast_to(lane, lanes.shape[:2] + (1,)), -1)[..., 0] >= dropout_threshold(p)
