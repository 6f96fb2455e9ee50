"""Drop-in for the reference's models/model.py: `make_model(opt)` / `TOD_ASR_Transformer_STC`.

Same constructor contract, forward signature, return tuple, parameter names and state_dict keys as the reference
(models/model.py:7-83 wrapping a HuggingFace BertModel / XLMRobertaModel and models/modules/hierarchical_classifier.py),
but every device operation runs in the hand-written sm_100a kernels of libnbest_sm100.so:

  * inputs are packed on the GPU (cu_seqlens + per-token maps), both streams (ASR n-best and transcript) go through
    the encoder as ONE packed batch, padding does no work;
  * QKV / output / FFN projections are tcgen05 GEMMs with fused bias / GELU / dropout+residual epilogues; attention is
    a fused varlen kernel; embedding+LN, LN and their backward are fused vectorised kernels;
  * the head (11 Linear + sigmoid + 10 softmax + scatter + decode) and the 4-term loss are fused kernels;
  * parameters live in one flat fp32 master buffer (+ bf16 working copy, + flat fp32 gradient buffer) so that the
    fused BertAdam and the bucketed NCCL all-reduce work on contiguous memory.

There is no CPU / eager fallback: constructing the model without an sm_100 GPU raises.
"""
import os
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .optim import FlatBuffers, register_flat

H = 768  # reference models/model.py:30 hard-codes fea_dim = 768


@dataclass
class EncoderSpec:
    kind: str = "bert"            # "bert" | "roberta" | "xlm-roberta"  (reference MODEL_CLASSES, n_best_asr_bert.py:33-37)
    vocab_size: int = 30522
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    intermediate: int = 3072
    max_position: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12
    pad_token_id: int = 0
    hidden_dropout: float = 0.1   # HF hidden_dropout_prob (embeddings + both dense outputs)
    attn_dropout: float = 0.1     # HF attention_probs_dropout_prob

    @staticmethod
    def bert_base(**kw):
        return EncoderSpec(**kw)

    @staticmethod
    def xlmr_base(**kw):
        d = dict(kind="xlm-roberta", vocab_size=250002, max_position=514, type_vocab=1, ln_eps=1e-5, pad_token_id=1)
        d.update(kw)
        return EncoderSpec(**d)

    @staticmethod
    def roberta_base(**kw):
        d = dict(kind="roberta", vocab_size=50265, max_position=514, type_vocab=1, ln_eps=1e-5, pad_token_id=1)
        d.update(kw)
        return EncoderSpec(**d)

    @property
    def roberta_style(self):
        """RoBERTa-family embeddings: position ids = padding_idx + 1 + running count of non-pad tokens, <s> = 0 is
        masked as a key by the reference's `input_ids > 0` while <pad> = 1 stays attendable."""
        return self.kind != "bert"

    @staticmethod
    def from_hf(encoder):
        c = encoder.config
        mt = getattr(c, "model_type", "bert")
        kind = "xlm-roberta" if mt == "xlm-roberta" else ("roberta" if "roberta" in mt else "bert")
        return EncoderSpec(kind=kind, vocab_size=c.vocab_size, hidden=c.hidden_size, layers=c.num_hidden_layers,
                           heads=c.num_attention_heads, intermediate=c.intermediate_size,
                           max_position=c.max_position_embeddings, type_vocab=c.type_vocab_size, ln_eps=c.layer_norm_eps,
                           pad_token_id=c.pad_token_id if c.pad_token_id is not None else 0,
                           hidden_dropout=c.hidden_dropout_prob, attn_dropout=c.attention_probs_dropout_prob)


class _Tree(nn.Module):
    """Container used to reproduce the reference's dotted parameter names (bert_encoder.encoder.layer.0. ...)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("container module")


def _register(root, dotted, param):
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, _Tree())
        m = m._modules[p]
    m.register_parameter(parts[-1], param)


def _layer_names(l):
    p = "bert_encoder.encoder.layer.%d." % l
    return dict(qw=p + "attention.self.query.weight", qb=p + "attention.self.query.bias",
                kw=p + "attention.self.key.weight", kb=p + "attention.self.key.bias",
                vw=p + "attention.self.value.weight", vb=p + "attention.self.value.bias",
                aow=p + "attention.output.dense.weight", aob=p + "attention.output.dense.bias",
                g1=p + "attention.output.LayerNorm.weight", b1=p + "attention.output.LayerNorm.bias",
                iw=p + "intermediate.dense.weight", ib=p + "intermediate.dense.bias",
                ow=p + "output.dense.weight", ob=p + "output.dense.bias",
                g2=p + "output.LayerNorm.weight", b2=p + "output.LayerNorm.bias")


def _describe(t):
    return "%s %s %s" % (t.device, t.dtype, tuple(t.shape)) if torch.is_tensor(t) else type(t).__name__


class _Saved:
    """Activations one forward keeps for its backward."""
    pass


class TOD_ASR_Transformer_STC(nn.Module):
    """TOD ASR Transformer Semantic Tuple Classifier (reference models/model.py:11), B200-native."""

    def __init__(self, opt=None, spec=None, top2bottom=None, dropout=None, device=None, encoder_state=None,
                 none_bottoms=(), seed=999):
        super().__init__()
        if opt is not None:
            enc = getattr(opt, "pretrained_model", None)
            if spec is None:
                spec = enc if isinstance(enc, EncoderSpec) else EncoderSpec.from_hf(enc)
            if encoder_state is None and enc is not None and not isinstance(enc, EncoderSpec):
                encoder_state = {k: v.detach() for k, v in enc.state_dict().items()}
            top2bottom = opt.top2bottom_dict if top2bottom is None else top2bottom
            dropout = opt.dropout if dropout is None else dropout
            device = opt.device if device is None else device
            assert getattr(opt, "cls_type", "stc") == "stc"
            self.score_util = getattr(opt, "score_util", None)
            self.sent_repr = getattr(opt, "sent_repr", None)
            if getattr(opt, "label_vocab_size", None) is not None:
                assert opt.label_vocab_size == sum(len(v) for v in top2bottom.values())
        self.cls_type = "stc"
        device = torch.device(device if device is not None else "cuda")
        if device.type != "cuda":
            raise RuntimeError("nbest_b200 has no CPU path: opt.device must be a CUDA (sm_100) device")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if spec.hidden != H or spec.hidden // spec.heads != 64:
            raise ValueError("the hot path is built for hidden = 768, head_dim = 64 (reference fea_dim, models/model.py:30)")
        self.spec = spec
        self.device = device
        self.head_dropout = float(dropout if dropout is not None else 0.0)
        self.hier = ops.DeviceHierarchy(top2bottom, none_bottoms, device=device)
        self.top2bottom_dict = self.hier.top2bottom
        self._step_seed = int(seed)
        self.defer_low_wgrad = int(os.environ.get("NBEST_DEFER_WGRAD", "1"))     # lowest layers whose wgrads are deferred, see _encoder_backward
        self._seed_salt = 0      # data-parallel rank salt (trainer): ranks draw independent dropout masks
        # The head (and the MSE term) only ever read the [CLS] row of the last hidden state (reference models/model.py:
        # 46-47,58), so the last layer's attention output, out-projection, FFN and LayerNorms are computed for that row
        # only: ~1/12 less encoder work, identical results (with hidden dropout the compact rows draw a different, equally
        # valid mask: mask indices follow the row layout). False = run the last layer on every token.
        self.cls_only_last_layer = True
        # Attention of sequences <= 128 tokens on the tcgen05 / TMEM / TMA tile kernels (csrc/attention_tc.cu); longer ones
        # (10-best inference) keep the block-loop kernels of csrc/attention.cu. NBEST_ATTN_TC=0: everything on the latter.
        self.attn_tensor_path = os.environ.get("NBEST_ATTN_TC", "1") != "0"
        self._build_params(encoder_state, seed)

    # ------------------------------------------------------------------------------------------------ parameters
    def _build_params(self, encoder_state, seed):
        s, hier, dev = self.spec, self.hier, self.device
        I = s.intermediate
        e = "bert_encoder.embeddings."
        reg = OrderedDict()          # registration (reference / HF) order: name -> shape
        reg[e + "word_embeddings.weight"] = (s.vocab_size, H)
        reg[e + "position_embeddings.weight"] = (s.max_position, H)
        reg[e + "token_type_embeddings.weight"] = (s.type_vocab, H)
        reg[e + "LayerNorm.weight"] = (H,)
        reg[e + "LayerNorm.bias"] = (H,)
        flat_order = list(reg.keys())
        aligns = [64] * len(flat_order)
        for l in range(s.layers):
            n = _layer_names(l)
            for k, shp in (("qw", (H, H)), ("qb", (H,)), ("kw", (H, H)), ("kb", (H,)), ("vw", (H, H)), ("vb", (H,)),
                           ("aow", (H, H)), ("aob", (H,)), ("g1", (H,)), ("b1", (H,)), ("iw", (I, H)), ("ib", (I,)),
                           ("ow", (H, I)), ("ob", (H,)), ("g2", (H,)), ("b2", (H,))):
                reg[n[k]] = shp
            # flat order keeps q|k|v weights and q|k|v biases adjacent: fused [2304,768] / [2304] views
            flat_order += [n[k] for k in ("qw", "kw", "vw", "qb", "kb", "vb", "aow", "aob", "g1", "b1", "iw", "ib", "ow", "ob", "g2", "b2")]
            aligns += [64] * 16
        reg["bert_encoder.pooler.dense.weight"] = (H, H)
        reg["bert_encoder.pooler.dense.bias"] = (H,)
        flat_order += ["bert_encoder.pooler.dense.weight", "bert_encoder.pooler.dense.bias"]
        aligns += [64, 64]
        head_w = ["clf.top_linear_layer.weight"] + ["clf.linear_layers.lin_%d.weight" % k for k in hier.group_tops]
        head_b = ["clf.top_linear_layer.bias"] + ["clf.linear_layers.lin_%d.bias" % k for k in hier.group_tops]
        sizes = [hier.n_top] + [len(hier.top2bottom[k]) for k in hier.group_tops]
        for wn, bn, n in zip(head_w, head_b, sizes):
            reg[wn] = (n, H)
            reg[bn] = (n,)
        flat_order += head_w + head_b
        aligns += [64] * len(head_w) + [64] + [1] * (len(head_b) - 1)    # biases packed: one contiguous [n_cols] vector

        self._names = flat_order
        self._index = {n: i for i, n in enumerate(flat_order)}
        self.flat = FlatBuffers([reg[n] for n in flat_order], dev, with_bf16=True, aligns=aligns)
        f = self.flat
        # ---- initial values
        g = torch.Generator(device=dev).manual_seed(int(seed))
        with torch.no_grad():
            for i, n in enumerate(flat_order):
                v = f.view(f.params, i)
                if encoder_state is not None and n.startswith("bert_encoder."):
                    key = n[len("bert_encoder."):]
                    if key in encoder_state:
                        v.copy_(encoder_state[key].to(dev, torch.float32))
                        continue
                    if "pooler" in n:              # add_pooling_layer=False encoders: keep the key, zero weights
                        continue
                    raise KeyError("encoder state is missing %s" % key)
                if n.startswith("clf."):           # nn.Linear default init (init_weight is never called)
                    bound = 1.0 / np.sqrt(H)
                    v.copy_((torch.rand(v.shape, device=dev, generator=g) * 2 - 1) * bound)
                elif n.endswith("LayerNorm.weight"):
                    v.fill_(1.0)
                elif n.endswith(".bias"):
                    v.zero_()
                else:
                    v.copy_(torch.randn(v.shape, device=dev, generator=g) * 0.02)
            if encoder_state is None:
                f.view(f.params, self._index[e + "word_embeddings.weight"])[s.pad_token_id].zero_()
                if s.roberta_style:
                    f.view(f.params, self._index[e + "position_embeddings.weight"])[1].zero_()
        # ---- nn.Parameters are views of the flat master buffer, .grad views of the flat gradient buffer
        plist = []
        for n in reg:                               # reference order for named_parameters()
            i = self._index[n]
            p = nn.Parameter(f.view(f.params, i), requires_grad=True)
            p.grad = None if "pooler" in n else f.view(f.grads, i)   # pooler grads stay None (never used)
            _register(self, n, p)
        for n in flat_order:
            plist.append(self.get_parameter(n))
        self._plist = plist
        register_flat(f, plist)
        # ---- fused views
        def fused(buf, first, rows, cols=None):
            o = f.offsets[self._index[first]]
            n = rows * (cols or 1)
            t = buf[o:o + n]
            return t.view(rows, cols) if cols else t
        self._w = []
        for l in range(s.layers):
            n = _layer_names(l)
            d = {}
            for tag, buf in (("p", f.params), ("g", f.grads), ("h", f.bf16)):
                d[tag + "_wqkv"] = fused(buf, n["qw"], 3 * H, H)
                d[tag + "_bqkv"] = fused(buf, n["qb"], 3 * H)
                d[tag + "_wo"] = fused(buf, n["aow"], H, H)
                d[tag + "_bo"] = fused(buf, n["aob"], H)
                d[tag + "_g1"] = fused(buf, n["g1"], H)
                d[tag + "_b1"] = fused(buf, n["b1"], H)
                d[tag + "_w1"] = fused(buf, n["iw"], I, H)
                d[tag + "_bi"] = fused(buf, n["ib"], I)
                d[tag + "_w2"] = fused(buf, n["ow"], H, I)
                d[tag + "_b2o"] = fused(buf, n["ob"], H)
                d[tag + "_g2"] = fused(buf, n["g2"], H)
                d[tag + "_b2"] = fused(buf, n["b2"], H)
            self._w.append(d)
        self._emb = {}
        for tag, buf in (("p", f.params), ("g", f.grads)):
            self._emb[tag + "_word"] = fused(buf, e + "word_embeddings.weight", s.vocab_size, H)
            self._emb[tag + "_pos"] = fused(buf, e + "position_embeddings.weight", s.max_position, H)
            self._emb[tag + "_type"] = fused(buf, e + "token_type_embeddings.weight", s.type_vocab, H)
            self._emb[tag + "_gamma"] = fused(buf, e + "LayerNorm.weight", H)
            self._emb[tag + "_beta"] = fused(buf, e + "LayerNorm.bias", H)
        self._head = {}
        for tag, buf in (("p", f.params), ("g", f.grads)):
            self._head[tag + "_w"] = fused(buf, head_w[0], hier.n_cols, H)
            self._head[tag + "_b"] = fused(buf, head_b[0], hier.n_cols)

    # nn.Module plumbing -------------------------------------------------------------------------------------
    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, device=self.device))
        if probe.device != self.device or probe.dtype != torch.float32:
            raise RuntimeError("TOD_ASR_Transformer_STC lives in flat fp32 buffers on %s and cannot be moved or cast" % self.device)
        return self

    # non-parameter buffers that checkpoints written under older transformers versions carry (persistent there)
    _IGNORED_BUFFERS = ("embeddings.position_ids", "embeddings.token_type_ids")

    def load_state_dict(self, state_dict, strict=True, assign=False):
        missing = [n for n in self._names if n not in state_dict]
        unexpected = [k for k in state_dict if k not in self._index and not k.endswith(self._IGNORED_BUFFERS)]
        if strict and (missing or unexpected):
            raise RuntimeError("load_state_dict: missing %s unexpected %s" % (missing, unexpected))
        with torch.no_grad():
            for n in self._names:
                if n in state_dict:
                    self.flat.view(self.flat.params, self._index[n]).copy_(state_dict[n].to(self.device, torch.float32))
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def load_model(self, load_dir):
        """reference models/model.py:75-80"""
        self.load_state_dict(torch.load(open(load_dir, "rb"), map_location=self.device))

    def save_model(self, save_dir):
        """reference models/model.py:82-83"""
        torch.save(self.state_dict(), open(save_dir, "wb"))

    def state_dict(self, *args, **kwargs):
        """Independent fp32 tensors under the reference's keys (not views of the flat buffer)."""
        sd = super().state_dict(*args, **kwargs)
        return OrderedDict((k, v.detach().clone()) for k, v in sd.items())

    def _refresh_bf16(self):
        """bf16 working copy of the weights. nbest_b200.BertAdam refreshes it inside its update kernel; any torch-side
        in-place change of a parameter bumps the flat buffer's version counter and triggers one cast kernel here."""
        f = self.flat
        if getattr(f, "bf16_version", -1) != f.params._version:
            ops.cast_f32_bf16(f.params, f.bf16)
            f.bf16_version = f.params._version

    def zero_grad(self, set_to_none=False):
        ops.zero_(self.flat.grads)

    # ------------------------------------------------------------------------------------------------ packing
    def _pack_streams(self, input_ids, seg_ids, trans_input_ids, trans_seg_ids, lens=None, trans_lens=None,
                      drop_token_types=None):
        """Both streams -> ONE packed batch: ASR sequences first (gradient-carrying prefix), then transcripts."""
        kind = "xlm-roberta" if self.spec.roberta_style else "bert"      # packing mode (position ids / key mask)
        if drop_token_types is None:
            drop_token_types = self.spec.kind == "xlm-roberta"
        if drop_token_types:
            seg_ids = trans_seg_ids = None                      # reference models/model.py:42-43: no token types
        elif self.spec.type_vocab < 2:
            for t in (seg_ids, trans_seg_ids):                  # HF would raise IndexError in the token-type embedding
                if t is not None and int(t.max()) >= self.spec.type_vocab:
                    raise IndexError("token type id %d out of range for type_vocab_size %d" % (int(t.max()), self.spec.type_vocab))
        if trans_input_ids is not None:
            # one packing pass writes both streams in place (no concatenation kernels)
            pk = ops.pack_batch_dual(input_ids, seg_ids, lens, trans_input_ids, trans_seg_ids, trans_lens, kind)
        else:
            pa = ops.pack_batch(input_ids, seg_ids, kind, lens)
        if trans_input_ids is None:
            pa.B_asr, pa.T_asr, pa.max_len_asr = pa.B, pa.T, pa.max_len
            pa.sum_l2_asr = pa.sum_l2
            if self.attn_tensor_path:
                pa.plan = ops.attn_plan(pa.cu_seqlens, pa.seq_of, pa.B, pa.T)
            return pa
        pk.plan = ops.attn_plan(pk.cu_seqlens, pk.seq_of, pk.B, pk.T, break_at=pk.B_asr) if self.attn_tensor_path else None
        return pk

    def _seed(self, layer, site):
        return ((self._step_seed * 1000003 + layer * 16 + site) ^ self._seed_salt) & 0xFFFFFFFF

    # ------------------------------------------------------------------------------------------------ encoder fwd
    def _encode(self, pk, save):
        s, dev = self.spec, self.device
        T = pk.T
        train = self.training
        p_h = s.hidden_dropout if train else 0.0
        p_a = s.attn_dropout if train else 0.0
        bf = lambda *shape: torch.empty(shape, device=dev, dtype=torch.bfloat16)
        f32 = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        self._refresh_bf16()
        sv = _Saved()
        sv.pk, sv.p_h, sv.p_a, sv.layers = pk, p_h, p_a, []
        em = self._emb
        x = bf(T, H)
        sv.mean0, sv.rstd0 = f32(T), f32(T)
        ops.embed_ln_fwd(pk, em["p_word"], em["p_pos"], em["p_type"], em["p_gamma"], em["p_beta"], s.ln_eps, x, sv.mean0,
                         sv.rstd0, p_h, self._seed(0, 15))
        kv = pk.key_valid if s.roberta_style else None             # BERT: every in-sequence key is valid (ids > 0)
        qkv = ctx = pre1 = x1 = u = gact = pre2 = None
        cls_last = self.cls_only_last_layer and s.layers >= 1
        sv.cls_compact = cls_last
        for l in range(s.layers):
            w = self._w[l]
            L = _Saved()
            last_cls = cls_last and l == s.layers - 1
            # rows that the post-attention block (out-proj .. LN2) works on: every token, or one [CLS] row per sequence in
            # the last layer (only that row is consumed downstream: reference models/model.py:46-47,58)
            n = pk.B if last_cls else T
            if save or qkv is None:
                qkv = bf(T, 3 * H)
            if save or ctx is None or ctx.shape[0] != n:
                ctx, pre1, x1 = bf(n, H), bf(n, H), bf(n, H)
                u = bf(n, s.intermediate) if save else None
                gact, pre2 = bf(n, s.intermediate), bf(n, H)
            L.x_in, L.n = x, n
            ops.gemm(x, w["h_wqkv"], epilogue=ops.EPI_BIAS, bias=w["p_bqkv"], out=qkv)
            if last_cls:
                L.lse = f32(s.heads, pk.B)
                ops.attn_cls_fwd(qkv, pk.cu_seqlens, kv, pk.B, pk.max_len, s.heads, T, ctx, L.lse, p_a, self._seed(l, 1))
                resid = ops.rows_gather(x, pk.cu_seqlens, pk.B, bf(pk.B, H))
            else:
                L.lse = f32(s.heads, T)
                if pk.plan is not None:
                    ops.attn_tiles_fwd(qkv, pk.plan, 0, kv, s.heads, T, ctx, L.lse, p_a, self._seed(l, 1), sum_l2=pk.sum_l2)
                    if pk.max_len > 128:
                        ops.attn_fwd(qkv, pk.cu_seqlens, kv, pk.B, pk.max_len, s.heads, T, ctx, L.lse, p_a, self._seed(l, 1),
                                     min_len=129)
                else:
                    ops.attn_fwd(qkv, pk.cu_seqlens, kv, pk.B, pk.max_len, s.heads, T, ctx, L.lse, p_a, self._seed(l, 1),
                                 sum_l2=pk.sum_l2)
                resid = x
            # the GEMM epilogue also emits the rows' partial {sum, sum of squares}: the LayerNorm behind it is single-pass
            part = self._row_partials(n)
            ops.gemm(ctx, w["h_wo"], epilogue=ops.EPI_BIAS_DROP_RES, bias=w["p_bo"], aux=resid, out=pre1, out2=part, p_drop=p_h,
                     seed=self._seed(l, 2))
            L.mean1, L.rstd1 = f32(n), f32(n)
            ops.ln_fwd(pre1, w["p_g1"], w["p_b1"], s.ln_eps, x1, L.mean1, L.rstd1, row_partials=part)
            # gact = gelu(x1 W1 + b); u = gelu'(x1 W1 + b) — the derivative is saved, so the backward epilogue only multiplies
            ops.gemm(x1, w["h_w1"], epilogue=ops.EPI_BIAS_GELU, bias=w["p_bi"], out=gact, out2=u)
            ops.gemm(gact, w["h_w2"], epilogue=ops.EPI_BIAS_DROP_RES, bias=w["p_b2o"], aux=x1, out=pre2, out2=part, p_drop=p_h,
                     seed=self._seed(l, 3))
            L.mean2, L.rstd2 = f32(n), f32(n)
            x_out = bf(n, H) if (save or last_cls) else self._pingpong(x, T)
            ops.ln_fwd(pre2, w["p_g2"], w["p_b2"], s.ln_eps, x_out, L.mean2, L.rstd2, row_partials=part)
            L.qkv, L.ctx, L.pre1, L.x1, L.u, L.g, L.pre2 = qkv, ctx, pre1, x1, u, gact, pre2
            if save:
                sv.layers.append(L)
            x = x_out
        sv.x_last = x            # [T,768], or [B,768] (row b = sequence b) when the last layer ran on the CLS rows only
        return sv

    def _cls_cu(self, sv, row0, B):
        """cu_seqlens-style row index of the CLS vectors of sequences row0 .. row0+B inside sv.x_last."""
        if sv.cls_compact:
            ar = getattr(self, "_arange_i32", None)
            if ar is None or ar.numel() < sv.pk.B + 1:
                ar = torch.arange(max(1024, sv.pk.B + 1), device=self.device, dtype=torch.int32)
                self._arange_i32 = ar
            return ar[row0:row0 + B + 1]
        return sv.pk.cu_seqlens[row0:row0 + B + 1]

    def _row_partials(self, n):
        """fp32 [n, 12, 2] workspace for the LayerNorm partial statistics a DROP_RES GEMM epilogue writes (re-used: each
        GEMM -> LayerNorm pair is adjacent in stream order)."""
        if not getattr(self, "ln_fused_stats", True):
            return None
        ws = getattr(self, "_part_ws", None)
        if ws is None or ws.shape[0] < n:
            ws = torch.empty((max(n, 1024), H // 64, 2), device=self.device, dtype=torch.float32)
            self._part_ws = ws
        return ws[:n]

    def _pingpong(self, x, T):
        """Inference: two alternating hidden-state buffers instead of one per layer."""
        bufs = getattr(self, "_pp", None)
        if bufs is None or bufs[0].shape[0] < T or bufs[0].device != x.device:
            bufs = [torch.empty((T, H), device=self.device, dtype=torch.bfloat16) for _ in range(2)]
            self._pp = bufs
        a, b = bufs[0][:T], bufs[1][:T]
        return b if x.data_ptr() == a.data_ptr() else a

    # ------------------------------------------------------------------------------------------------ encoder bwd
    def _encoder_backward(self, sv, dx, T_act, B_act):
        """dx: [T_act, 768] bf16 gradient w.r.t. the last hidden state of the first T_act tokens (B_act sequences)."""
        s, dev = self.spec, self.device
        pk, p_h, p_a = sv.pk, sv.p_h, sv.p_a
        T = pk.T
        self._last_pk, self._last_T_act = pk, T_act      # (the trainer's row-sparse embedding-gradient exchange reads the token list)
        bf = lambda *shape: torch.empty(shape, device=dev, dtype=torch.bfloat16)
        kv = pk.key_valid if s.roberta_style else None
        cu = pk.cu_seqlens[:B_act + 1]
        max_len = pk.max_len_asr if B_act == pk.B_asr and B_act != pk.B else pk.max_len
        dqkv = bf(T_act, 3 * H)
        delta = torch.empty((s.heads, T_act), device=dev, dtype=torch.float32)     # written by the out-proj dgrad (EPI_DELTA)
        A = lambda t: t[:T_act]
        ws = {}       # per-row-count workspaces of the post-attention block: {n: (dpre, dprem, du, dx1, dctx)}
        # The weight-gradient GEMMs of the LOWEST layer(s) are issued after the embedding backward instead of inside the
        # layer: nothing downstream needs them, and in a data-parallel run the embedding bucket — the largest one, and the
        # last the backward can finish — then has its all-reduce and update running under ~0.2 ms of GEMMs per deferred layer
        # instead of after the last kernel of the step (profiles/r2_dp_timeline_*.log). Same kernels, same operands:
        # results are unchanged; a deferred layer keeps its own dm / du / dqkv buffers alive until then.
        deferred = []         # [(layer, dqkv, [(args, kwargs) of its wgrad GEMMs])], highest layer first
        for l in reversed(range(s.layers)):
            w, L = self._w[l], sv.layers[l]
            compact = sv.cls_compact and l == s.layers - 1       # this layer's tail ran on one CLS row per sequence
            n = B_act if compact else T_act
            if n not in ws:
                ws[n] = (bf(n, H), (bf(n, H) if p_h > 0 else None), bf(n, s.intermediate), bf(n, H), bf(n, H))
            dpre, dprem, du, dx1, dctx = ws[n]
            defer = l < self.defer_low_wgrad and not compact
            if defer:
                dpre, dprem, du = bf(n, H), (bf(n, H) if p_h > 0 else None), bf(n, s.intermediate)
                dqkv = bf(T_act, 3 * H)
                deferred.append((l, dqkv, []))
            wgrad = (lambda *a, **k: deferred[-1][2].append((a, k))) if defer else ops.gemm
            R = lambda t: t[:n]
            # ---- FFN block
            ops.ln_bwd(dx, R(L.pre2), L.mean2, L.rstd2, w["p_g2"], dpre, w["g_g2"], w["g_b2"], dx_masked=dprem,
                       dbias=w["g_b2o"], p_drop=p_h, seed=self._seed(l, 3), T=n)
            dm = dprem if p_h > 0 else dpre
            # du = (dm W2) * gelu'(u), and in the same epilogue the FFN-in bias gradient g_bi += column sums of du
            ops.gemm(dm, w["h_w2"], b_mn_major=True, epilogue=ops.EPI_DGELU, aux=R(L.u), out=du, out2=w["g_bi"])
            wgrad(dm, R(L.g), a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_ACCUM_F32, out=w["g_w2"])
            ops.gemm(du, w["h_w1"], b_mn_major=True, epilogue=ops.EPI_ADD, aux=dpre, out=dx1)              # + residual grad
            wgrad(du, R(L.x1), a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_ACCUM_F32, out=w["g_w1"])
            if defer:        # the FFN block's dm stays alive for its deferred wgrad: the attention block gets buffers of its own
                dpre, dprem = bf(n, H), (bf(n, H) if p_h > 0 else None)
            # ---- attention block
            ops.ln_bwd(dx1, R(L.pre1), L.mean1, L.rstd1, w["p_g1"], dpre, w["g_g1"], w["g_b1"], dx_masked=dprem,
                       dbias=w["g_bo"], p_drop=p_h, seed=self._seed(l, 2), T=n)
            dm = dprem if p_h > 0 else dpre
            if compact:
                ops.gemm(dm, w["h_wo"], b_mn_major=True, epilogue=ops.EPI_NONE, out=dctx)
            else:      # dO = dm Wo, and in the same epilogue the attention backward's delta = rowsum(dO * O) per head
                ops.gemm(dm, w["h_wo"], b_mn_major=True, epilogue=ops.EPI_DELTA, aux=R(L.ctx), out=dctx, out2=delta[:, :n])
            wgrad(dm, R(L.ctx), a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_ACCUM_F32, out=w["g_wo"])
            # a window of short, non-persistent kernels (attention backward): the data-parallel trainer starts the pending
            # gradient all-reduces here, so that they run next to kernels that shrink gracefully instead of next to the
            # persistent GEMMs, whose CTA pairs would have to wait for the SMs the collective occupies
            self._notify("slot")
            if compact:
                ops.attn_cls_bwd(L.qkv, cu, kv, B_act, max_len, s.heads, T, L.ctx, dctx, L.lse, pk.B, dqkv, p_a, self._seed(l, 1))
                # the residual branch reaches the layer input only at the CLS rows
                dres = ops.rows_scatter(dpre, cu, B_act, T_act, bf(T_act, H))
                dx = bf(T_act, H)
            else:
                l2s = pk.sum_l2 if B_act == pk.B else pk.sum_l2_asr
                if pk.plan is not None:
                    ops.attn_tiles_bwd(L.qkv, pk.plan, 0 if B_act == pk.B else 1, kv, s.heads, T, T_act, dctx, L.lse, delta, T_act,
                                       dqkv, p_a, self._seed(l, 1), sum_l2=l2s)
                    if max_len > 128:
                        ops.attn_bwd(L.qkv, cu, kv, B_act, max_len, s.heads, T, None, dctx, L.lse, dqkv, delta, p_a,
                                     self._seed(l, 1), T_active=T_act, min_len=129)
                else:
                    ops.attn_bwd(L.qkv, cu, kv, B_act, max_len, s.heads, T, None, dctx, L.lse, dqkv, delta, p_a, self._seed(l, 1),
                                 T_active=T_act, sum_l2=l2s)
                dres = dpre
            ops.gemm(dqkv, w["h_wqkv"], b_mn_major=True, epilogue=ops.EPI_ADD, aux=dres, out=dx)   # (full path: dx was consumed above)
            wgrad(dqkv, A(L.x_in), a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_ACCUM_F32, out=w["g_wqkv"])
            if not defer:
                ops.colsum(dqkv, w["g_bqkv"], T=T_act)
                self._notify("layer%d" % l)
        em = self._emb
        ops.embed_ln_bwd(pk, em["p_word"], em["p_pos"], em["p_type"], em["p_gamma"], sv.mean0, sv.rstd0, dx, em["g_word"],
                         em["g_pos"], em["g_type"], em["g_gamma"], em["g_beta"], p_h, self._seed(0, 15),
                         word_pad_row=s.pad_token_id, pos_pad_row=1 if s.roberta_style else -1, T=T_act)
        self._notify("emb")
        for l, dqkv_l, calls in deferred:
            for a, k in calls:
                ops.gemm(*a, **k)
            ops.colsum(dqkv_l, self._w[l]["g_bqkv"], T=T_act)
            self._notify("layer%d" % l)

    def _notify(self, bucket):
        """Tell the data-parallel trainer that every gradient kernel of `bucket` has been enqueued."""
        hook = getattr(self, "_grad_ready_hook", None)
        if hook is not None:
            hook(bucket)

    # ------------------------------------------------------------------------------------------------ head
    def _head_forward(self, sv, B, row0=0):
        hier, dev = self.hier, self.device
        f32 = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        o = _Saved()
        o.cls, o.logits = f32(B, H), f32(B, hier.n_cols)
        o.top, o.bottom, o.final = f32(B, hier.n_top), f32(B, hier.n_cols - hier.n_top), f32(B, hier.n_bottom)
        o.decode = torch.empty((B, hier.n_bottom), device=dev, dtype=torch.uint8)
        o.p = self.head_dropout if self.training else 0.0
        o.seed = self._seed(99, 7)
        ops.stc_head_fwd(sv.x_last, self._cls_cu(sv, row0, B), B, self._head["p_w"], self._head["p_b"], hier, o.cls,
                         o.logits, o.top, o.bottom, o.final, o.decode, o.p, o.seed)
        return o

    def _cls_rows(self, sv, row0, B):
        out = torch.empty((B, H), device=self.device, dtype=torch.float32)
        return ops.rows_gather(sv.x_last, self._cls_cu(sv, row0, B), B, out)

    def _backward_from_dlogits(self, sv, ho, dlogits, d_cls_asr, d_cls_trans, head_on_trans=False):
        """dlogits [B,n_cols] (+ optional direct gradients of the two CLS vectors) -> all parameter gradients."""
        pk, dev = sv.pk, self.device
        B = pk.B_asr
        dcls_head = torch.empty((B, H), device=dev, dtype=torch.float32)
        ops.stc_head_bwd(dlogits, ho.cls, self._head["p_w"], self.hier, self._head["g_w"], self._head["g_b"], dcls_head,
                         accumulate_dcls=False, p_drop=ho.p, seed=ho.seed)
        self._notify("head")
        d_asr, d_trans = (None, dcls_head) if head_on_trans else (dcls_head, None)
        if d_cls_asr is not None:
            d_asr = d_cls_asr if d_asr is None else d_asr + d_cls_asr
        if d_cls_trans is not None:
            d_trans = d_cls_trans if d_trans is None else d_trans + d_cls_trans
        if d_trans is not None and pk.B > B:
            if d_asr is None:
                d_asr = torch.zeros((B, H), device=dev, dtype=torch.float32)
            dcls = torch.cat([d_asr, d_trans], 0)
            T_act, B_act = pk.T, pk.B
        else:
            if d_asr is None:
                return
            dcls, T_act, B_act = d_asr, pk.T_asr, B
        if sv.cls_compact:       # gradient of the compact [B,768] last hidden state: row b = sequence b
            dx = torch.empty((B_act, H), device=dev, dtype=torch.bfloat16)
            ops.cls_scatter(dcls.contiguous(), self._cls_cu(sv, 0, B_act), B_act, B_act, dx)
        else:
            dx = torch.empty((T_act, H), device=dev, dtype=torch.bfloat16)
            ops.cls_scatter(dcls.contiguous(), pk.cu_seqlens, B_act, T_act, dx)
        self._encoder_backward(sv, dx, T_act, B_act)

    # ------------------------------------------------------------------------------------------------ public: drop-in forward
    @ops.with_bound_stream
    def forward(self, opt, input_ids, trans_input_ids=None, seg_ids=None, trans_seg_ids=None, return_attns=False,
                classifier_input_type="asr", input_lens=None, trans_input_lens=None):
        """Reference signature (models/model.py:35). Returns (top_scores [B,30], {'lin_k': [B,n_k]}, final_scores [B,161],
        asr_cls [B,768], trans_cls [B,768] | None), autograd-connected: `.backward()` on anything computed from them runs
        the hand-written backward and accumulates into the parameters' .grad (views of the flat gradient buffer)."""
        if return_attns:
            raise NotImplementedError("return_attns=True is dead code in the reference (models/model.py:70-71 uses an undefined name)")
        need_grad = torch.is_grad_enabled()
        self._step_seed += 1
        # models/model.py:42-45: token types are dropped iff opt.pre_trained_model == "xlm-roberta" (the option string, not
        # the encoder class); a plain opt without the attribute falls back to the encoder kind
        ptm = getattr(opt, "pre_trained_model", None)
        drop_tt = (ptm == "xlm-roberta") if ptm else None
        pk = self._pack_streams(input_ids, seg_ids, trans_input_ids, trans_seg_ids, input_lens, trans_input_lens, drop_tt)
        sv = self._encode(pk, save=need_grad)
        on_trans = classifier_input_type == "transcript" and trans_input_ids is not None
        B = pk.B_asr
        ho = self._head_forward(sv, B, row0=B if on_trans else 0)
        if on_trans:
            trans_cls, asr_cls = ho.cls, self._cls_rows(sv, 0, B)
        else:
            asr_cls = ho.cls
            trans_cls = self._cls_rows(sv, B, B) if trans_input_ids is not None else None
        self.last_decode = ho.decode
        if need_grad:
            anchor = self._plist[0]
            outs = _STCFunction.apply(anchor, self, sv, ho, on_trans, ho.top, ho.bottom, ho.final, asr_cls,
                                      trans_cls if trans_cls is not None else torch.empty(0, device=self.device))
            top, bottom, final, asr_cls, tc = outs
            trans_cls = tc if trans_cls is not None else None
        else:
            top, bottom, final = ho.top, ho.bottom, ho.final
        bottoms = OrderedDict()
        for g, k in enumerate(self.hier.group_tops):
            c0, c1 = self.hier.grp_off_host[g] - self.hier.n_top, self.hier.grp_off_host[g + 1] - self.hier.n_top
            bottoms["lin_%d" % k] = bottom[:, c0:c1]
        return top, bottoms, final, asr_cls, trans_cls

    # ------------------------------------------------------------------------------------------------ public: fused step
    @ops.with_bound_stream
    def forward_loss_backward(self, input_ids, labels, trans_input_ids=None, seg_ids=None, trans_seg_ids=None,
                              add_l2_loss=False, mse_scale=1.0, input_lens=None, trans_input_lens=None, backward=True,
                              n_real=None):
        """Fused training path: model forward + cal_total_loss (n_best_asr_bert.py:160-195) + backward in our kernels.

        Returns (losses, head) where losses is a device fp32 tensor [mse, bce_final, bce_top, ce] (no host sync) and head
        carries top/bottom/final scores and the decode bitmap. total = losses.sum(); loss_record = total / B.
        mse_scale lets a data-parallel trainer scale the mean-reduced MSE term by 1/world_size.
        n_real: only the first n_real utterances are real; the rows behind them are FILLER sequences that quantise the
        batch's token counts for CUDA-graph replay (graph.add_fillers). They run through the encoder like any sequence
        but take no part in the loss: labels is [n_real, n_bottom], their loss gradient is exactly zero, and so is
        everything the backward accumulates for them."""
        nb = self.hier.n_bottom
        Br = input_ids.shape[0] if n_real is None else int(n_real)
        if not 0 < Br <= input_ids.shape[0]:
            raise ValueError("n_real=%d outside (0, B=%d]" % (Br, input_ids.shape[0]))
        if not (torch.is_tensor(labels) and labels.is_cuda and labels.dtype == torch.float32 and labels.dim() == 2
                and labels.shape == (Br, nb) and labels.is_contiguous()):
            raise ValueError("labels must be a contiguous CUDA float32 [B=%d, %d] multi-hot tensor (collate_fn, "
                             "tod_asr_util.py:118-130), got %s" % (Br, nb, _describe(labels)))
        self._step_seed += 1
        pk = self._pack_streams(input_ids, seg_ids, trans_input_ids, trans_seg_ids, input_lens, trans_input_lens)
        sv = self._encode(pk, save=backward)
        B = pk.B_asr
        ho = self._head_forward(sv, B)
        ho.n_real = Br
        dev = self.device
        losses = ops.zero_(torch.empty(4, device=dev, dtype=torch.float32))
        dlogits = torch.empty((B, self.hier.n_cols), device=dev, dtype=torch.float32)
        use_l2 = add_l2_loss and trans_input_ids is not None
        trans_cls = d_asr = d_trans = None
        if trans_input_ids is not None:
            if pk.B - B < B:
                raise ValueError("the transcript stream has fewer rows (%d) than the ASR stream (%d)" % (pk.B - B, B))
            trans_cls = self._cls_rows(sv, B, B)      # the reference always computes it (models/model.py:51-58)
        if use_l2:
            d_asr = torch.empty((B, H), device=dev, dtype=torch.float32)
            d_trans = torch.empty((B, H), device=dev, dtype=torch.float32)
        if Br < B:                                    # filler rows: zero loss gradient
            ops.zero_(dlogits[Br:])
            if use_l2:
                ops.zero_(d_asr[Br:])
                ops.zero_(d_trans[Br:])
        R = (lambda t: t[:Br]) if Br < B else (lambda t: t)
        ops.stc_loss_fwd_bwd(R(ho.logits), labels, self.hier, losses, R(dlogits), R(ho.cls) if use_l2 else None,
                             R(trans_cls) if use_l2 else None, mse_scale, R(d_asr) if use_l2 else None,
                             R(d_trans) if use_l2 else None)
        if backward:
            self._backward_from_dlogits(sv, ho, dlogits, d_asr, d_trans)
        ho.trans_cls = trans_cls
        self.last_decode = ho.decode
        return losses, ho

    @torch.no_grad()
    @ops.with_bound_stream
    def infer(self, input_ids, seg_ids=None, input_lens=None):
        """Inference step (eval_epoch, n_best_asr_bert.py:316-344): forward + decode bitmap, no activations kept."""
        was = self.training
        self.training = False
        try:
            pk = self._pack_streams(input_ids, seg_ids, None, None, input_lens, None)
            sv = self._encode(pk, save=False)
            return self._head_forward(sv, pk.B_asr)
        finally:
            self.training = was


class _STCFunction(torch.autograd.Function):
    """Connects the fused forward to autograd so that the reference's own cal_total_loss + total_loss.backward()
    (n_best_asr_bert.py:262-264) drive the hand-written backward."""

    @staticmethod
    def forward(ctx, anchor, model, sv, ho, on_trans, top, bottom, final, asr_cls, trans_cls):
        ctx.model, ctx.sv, ctx.ho, ctx.on_trans = model, sv, ho, on_trans
        ctx.set_materialize_grads(False)       # outputs the loss never touched (e.g. trans_cls) must arrive as None
        ctx.has_trans = trans_cls.numel() > 0
        ctx.save_for_backward(top, bottom)
        return top.view_as(top), bottom.view_as(bottom), final.view_as(final), asr_cls.view_as(asr_cls), trans_cls.view_as(trans_cls)

    @staticmethod
    @ops.with_bound_stream
    def backward(ctx, d_top, d_bottom, d_final, d_asr, d_trans):
        model, sv, ho = ctx.model, ctx.sv, ctx.ho
        top, bottom = ctx.saved_tensors
        B = top.shape[0]
        dlogits = torch.empty((B, model.hier.n_cols), device=model.device, dtype=torch.float32)
        c = lambda t: None if t is None else t.contiguous().float()
        ops.stc_scores_bwd(top, bottom, c(d_top), c(d_bottom), c(d_final), model.hier, dlogits)
        if not ctx.has_trans:
            d_trans = None
        model._backward_from_dlogits(sv, ho, dlogits, c(d_asr), c(d_trans), head_on_trans=ctx.on_trans)
        return (None,) * 10


def make_model(opt):
    """reference models/model.py:7-9"""
    return TOD_ASR_Transformer_STC(opt)
