"""CUDA-graph replay of the fused training step on one GPU.

The reference's step is ~3,300 eager launches issued from Python (n_best_asr_bert.py:254-277); ours is ~250 launches of
our own kernels, and at the reference's batch size (16, run/*.sh) those are so short that the step is bound by the host
issuing them (3.6 ms of Python + ctypes per step against ~1 ms of GPU work; at batch 256 the same host time hides behind
9.6 ms of kernels but shows up whenever the host has to wait for a result, i.e. in the end-to-end number). A captured
graph replays the whole step — packing, encoder forward, head, loss, backward, BertAdam, gradient zero-fill — with ONE
launch.

What a graph bakes in, and how each is handled:
  * by-value kernel arguments that change every step (dropout seeds, BertAdam's schedule multiplier): read from the
    context's device-side step state instead (include/nbest_sm100.h nbest_ctx_set_step_state), written by a one-thread
    kernel ahead of every replay;
  * device pointers: inputs are copied into static buffers owned by the graph's entry (H2D from pinned host memory, or
    D2D); activations live in the graphs' shared private memory pool;
  * shapes: the packed layout makes every kernel's shape a function of the TOKEN COUNTS of the two streams, which differ
    from batch to batch. `add_fillers` appends a few filler sequences to each stream so that the counts land on multiples
    of `multiple` tokens: batches then fall into a small set of shapes, one graph each (captured at first sight, LRU
    bounded). Fillers are ordinary sequences for the encoder and take no part in the loss (model.forward_loss_backward
    n_real): their loss gradient is exactly zero, so every gradient they contribute is an exact zero.

Scope: world size 1 and BertAdam (the reference's default --optim_choice); anything else, and any shape whose capture
fails, falls back to the eager DataParallelTrainer.step. No CPU path here either: capture needs the CUDA library.
"""
from collections import OrderedDict

import torch

from . import _lib, ops
from .optim import BertAdam, schedule_multiplier


def add_fillers(ids, seg, lens, n_fill=3, multiple=256, width=None, filler_id=1, max_fill_len=128):
    """Append `n_fill` filler rows to a right-padded [B, S] id tensor so that the stream's token count becomes a multiple
    of `multiple`. Returns (ids', seg', lens') with B + n_fill rows of width max(S, width, longest filler); `lens` is the
    host list of true lengths (prepare_inputs_for_roberta's third result). Every filler has 1..max_fill_len tokens of
    `filler_id` (any id > 0: BERT's mask is ids > 0, models/model.py:43), so n_fill * max_fill_len >= multiple + n_fill - 1
    is required."""
    lens = [int(x) for x in lens]
    B, S = ids.shape
    if n_fill * max_fill_len < multiple + n_fill - 1:
        raise ValueError("%d fillers of <= %d tokens cannot close a gap of up to %d tokens" % (n_fill, max_fill_len, multiple + n_fill - 1))
    T = sum(lens)
    target = -(-(T + n_fill) // multiple) * multiple
    rest = target - T                                   # n_fill <= rest < multiple + n_fill
    fill = []
    for i in range(n_fill):
        k = -(-rest // (n_fill - i))                    # as even as possible: the longest filler is ceil(rest / n_fill)
        fill.append(k)
        rest -= k
    W = max(S, int(width or 0), max(fill))
    out = ids.new_zeros((B + n_fill, W))
    out[:B, :S] = ids
    cols = torch.arange(W, device=ids.device)
    out[B:] = (cols[None, :] < torch.tensor(fill, device=ids.device)[:, None]).to(ids.dtype) * filler_id
    seg_out = None
    if seg is not None:
        seg_out = seg.new_zeros((B + n_fill, W))
        seg_out[:B, :S] = seg
    return out, seg_out, lens + fill


class _Entry:
    __slots__ = ("graph", "static", "losses", "head", "n_kernels", "tables_id")


class GraphedTrainer:
    """trainer = GraphedTrainer(DataParallelTrainer(model, optimizer))
    losses = trainer.step(ids, labels, trans_ids, seg, trans_seg, lens, trans_lens)        # same call as the eager trainer

    `ids` ... may live on the host (pinned: the copies into the graph's static inputs are then the step's H2D traffic) or
    on the device. `losses` and `last_head` are the graph's static outputs: consume them (or enqueue their consumer on
    the stream) before the next step. `bucket=(n_fill, multiple)` applies add_fillers to both streams first."""

    def __init__(self, trainer, max_graphs=32, bucket=None, width=None):
        if trainer.world != 1:
            raise ValueError("CUDA-graph replay covers the single-GPU step; data-parallel runs use the eager trainer")
        if type(trainer.optimizer) is not BertAdam:
            raise ValueError("CUDA-graph replay is built for BertAdam (the reference's default --optim_choice)")
        self.trainer, self.model, self.optimizer = trainer, trainer.model, trainer.optimizer
        if self.model.spec.roberta_style:
            raise ValueError("CUDA-graph replay needs host-known sequence lengths (BERT packing); XLM-R keeps its <pad> tokens")
        self.ctx = _lib.context(self.model.device.index)
        self.max_graphs, self.bucket, self.width = int(max_graphs), bucket, width
        self.entries = OrderedDict()
        self.pool = None
        self.stream = torch.cuda.Stream(device=self.model.device)
        self.replays = self.captures = self.eager_steps = 0
        self.failed = set()
        self.capture_error = None
        self.last_head = None
        self.last_kernels = 0
        self.kernels_total = 0        # kernels executed on the device through this trainer (replayed graph nodes included)

    # ------------------------------------------------------------------ per-step scalars
    def _salt(self):
        """Dropout salt of the NEXT step: a mixed function of the model's step counter (never 0 by construction)."""
        return ops._seed(((self.model._step_seed + 1) * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF) | 1

    def _sched(self):
        o = self.optimizer
        g0 = o._uniform()
        steps = {s for p, s in zip(o._plist, o._steps) if p.grad is not None}
        if len(steps) != 1:
            raise RuntimeError("BertAdam: per-tensor step counts diverged")
        return schedule_multiplier(steps.pop(), g0["t_total"], g0["warmup"], g0["schedule"])

    # ------------------------------------------------------------------ capture / replay
    def _key(self, ids, labels, trans_ids, seg, trans_seg, lens, trans_lens, n_real):
        la = [int(x) for x in lens]
        ml_a = max(la)
        key = [tuple(ids.shape), seg is not None, sum(la), ml_a if ml_a > 128 else 0, int(n_real), tuple(labels.shape),
               bool(self.trainer.add_l2_loss), bool(self.model.training)]
        if trans_ids is not None:
            lt = [int(x) for x in trans_lens]
            ml_t = max(lt)
            key += [tuple(trans_ids.shape), trans_seg is not None, sum(lt), ml_t if ml_t > 128 else 0]
        return tuple(key)

    def _capture(self, key, ids, labels, trans_ids, seg, trans_seg, lens, trans_lens, n_real):
        dev, m, o = self.model.device, self.model, self.optimizer
        e = _Entry()
        mk = lambda t: None if t is None else torch.empty(t.shape, dtype=t.dtype, device=dev)
        e.static = dict(ids=mk(ids), labels=mk(labels), trans_ids=mk(trans_ids), seg=mk(seg), trans_seg=mk(trans_seg))
        for k, t in (("ids", ids), ("labels", labels), ("trans_ids", trans_ids), ("seg", seg), ("trans_seg", trans_seg)):
            if t is not None:
                e.static[k].copy_(t, non_blocking=True)
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        seed0, steps0 = m._step_seed, list(o._steps)
        s = e.static
        g = torch.cuda.CUDAGraph()
        l0 = self.ctx.launches()
        self.ctx.set_step_indirect(True)
        try:
            with torch.cuda.graph(g, pool=self.pool, stream=self.stream, capture_error_mode="thread_local"):
                e.losses = self.trainer.step(s["ids"], s["labels"], s["trans_ids"], s["seg"], s["trans_seg"], lens, trans_lens,
                                             n_real=n_real)
                e.head = self.trainer.last_head
                # last node: back to the eager state (salt 0), so launches outside the graphs see their by-value seeds
                self.ctx.set_step_state(0, 1.0, 1.0, 1.0, torch.cuda.current_stream().cuda_stream)
        finally:
            self.ctx.set_step_indirect(False)
            m._step_seed, o._steps = seed0, steps0          # capturing executed nothing
        e.n_kernels = self.ctx.launches() - l0
        e.graph = g
        e.tables_id = id(o._tables)
        self.captures += 1
        return e

    def _eager(self, ids, labels, trans_ids, seg, trans_seg, lens, trans_lens, n_real):
        dev = self.model.device
        d = lambda t: None if t is None else t.to(dev, non_blocking=True)
        l0 = self.ctx.launches()
        losses = self.trainer.step(d(ids), d(labels), d(trans_ids), d(seg), d(trans_seg), lens, trans_lens, n_real=n_real)
        self.last_head = self.trainer.last_head
        self.last_kernels = self.ctx.launches() - l0
        self.kernels_total += self.last_kernels
        self.eager_steps += 1
        return losses

    def step(self, ids, labels, trans_ids=None, seg=None, trans_seg=None, lens=None, trans_lens=None, n_real=None):
        if self.bucket is not None:
            n_fill, multiple = self.bucket
            n_real = ids.shape[0] if n_real is None else n_real
            ids, seg, lens = add_fillers(ids, seg, lens, n_fill, multiple, self.width)
            if trans_ids is not None:
                trans_ids, trans_seg, trans_lens = add_fillers(trans_ids, trans_seg, trans_lens, n_fill, multiple, self.width)
        if n_real is None:
            n_real = ids.shape[0]
        if lens is None or (trans_ids is not None and trans_lens is None):
            raise ValueError("CUDA-graph replay needs the host length lists (input_lens of prepare_inputs_for_roberta)")
        o = self.optimizer
        # the very first step runs eagerly: it builds the optimizer's device tables, sets kernel attributes, sizes caches
        if o._tables is None or getattr(o.flat, "m", None) is None:
            return self._eager(ids, labels, trans_ids, seg, trans_seg, lens, trans_lens, n_real)
        key = self._key(ids, labels, trans_ids, seg, trans_seg, lens, trans_lens, n_real)
        if key in self.failed:
            return self._eager(ids, labels, trans_ids, seg, trans_seg, lens, trans_lens, n_real)
        e = self.entries.get(key)
        if e is not None and e.tables_id != id(o._tables):
            self.entries.clear()                    # the optimizer rebuilt its tables (lr / weight-decay change): pointers moved
            e = None
        fresh = e is None
        if fresh:
            while len(self.entries) >= self.max_graphs:
                self.entries.popitem(last=False)
            try:
                e = self._capture(key, ids, labels, trans_ids, seg, trans_seg, lens, trans_lens, n_real)
            except Exception as exc:             # e.g. a shape whose path synchronises; the eager step is always available
                self.failed.add(key)
                self.capture_error = repr(exc)
                torch.cuda.synchronize()
                return self._eager(ids, labels, trans_ids, seg, trans_seg, lens, trans_lens, n_real)
            self.entries[key] = e
        else:
            self.entries.move_to_end(key)
            for k, t in (("ids", ids), ("labels", labels), ("trans_ids", trans_ids), ("seg", seg), ("trans_seg", trans_seg)):
                if t is not None:
                    e.static[k].copy_(t, non_blocking=True)
        self.ctx.set_step_state(self._salt(), self._sched(), 1.0, 1.0, torch.cuda.current_stream().cuda_stream)
        e.graph.replay()
        # host-side bookkeeping of what the replayed step did on the device
        self.model._step_seed += 1
        for k, p in enumerate(o._plist):
            if p.grad is not None:
                o._steps[k] += 1
        self.replays += 1
        self.last_head = e.head
        self.trainer.last_head = e.head
        self.last_kernels = e.n_kernels + 1
        self.kernels_total += self.last_kernels
        return e.losses
