"""BertAdam drop-in (reference models/optimization.py:183-302) backed by the fused multi-tensor kernel.

Same constructor and semantics as the reference class — per-tensor clip_grad_norm_(p, max_grad_norm), Adam moments
without bias correction, e = 1e-6, decoupled weight decay, `lr * WarmupLinearSchedule(step / t_total)` with the first
update at lr 0, tensors with `grad is None` skipped — but the 221 one-tensor param groups of
n_best_asr_bert.py:535-550 are updated by two kernel launches over one flat fp32 buffer instead of ~3 k launches.
"""
import ctypes as C

import numpy as np
import torch

from . import ops
from ._lib import AdamTensor

ALIGN = 64  # elements; every tensor starts on a 256-byte boundary of the flat buffers


def warmup_linear(progress, warmup):
    """WarmupLinearSchedule.get_lr_ (models/optimization.py:168-171)."""
    if progress < warmup:
        return progress / warmup
    return max((progress - 1.0) / (warmup - 1.0), 0.0)


def schedule_multiplier(step, t_total, warmup, schedule="warmup_linear"):
    """_LRSchedule.get_lr (models/optimization.py:53-70) for the schedules n_best_asr_bert.py can select."""
    if t_total is None or t_total < 0 or schedule in (None, "none"):
        return 1.0
    progress = float(step) / float(t_total)
    w = max(float(warmup), 0.0)
    if schedule == "warmup_linear":
        return warmup_linear(progress, w)
    if schedule == "warmup_constant":
        return progress / w if progress < w else 1.0
    raise ValueError("Invalid schedule parameter: {}".format(schedule))


def build_adam_tables(spec, device, chunk=16384):
    """spec: list of dict(offset, numel, lr, weight_decay, active) -> device tables for nbest_bertadam_step."""
    n = len(spec)
    arr = (AdamTensor * n)()
    chunks = []
    for i, s in enumerate(spec):
        arr[i].offset = int(s["offset"])
        arr[i].numel = int(s["numel"])
        arr[i].lr = float(s["lr"])
        arr[i].weight_decay = float(s["weight_decay"])
        arr[i].active = 1 if s["active"] else 0
        if s["active"]:
            for start in range(0, int(s["numel"]), chunk):
                chunks.append((i, start, min(chunk, int(s["numel"]) - start)))
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return dict(tensors=torch.from_numpy(raw).to(device), n_tensors=n,
                chunks=torch.tensor(chunks, dtype=torch.int32, device=device).contiguous(), n_chunks=len(chunks),
                norms=torch.zeros(n, dtype=torch.float32, device=device))


class FlatBuffers:
    """One flat fp32 buffer each for parameters, gradients and the two Adam moments (+ optional bf16 working copy)."""

    def __init__(self, shapes, device, with_bf16=True, aligns=None):
        """aligns[i] (elements, default ALIGN) is the alignment of tensor i's first element; 1 packs it directly
        behind its predecessor (used to keep the 11 STC-head biases one contiguous [171] vector)."""
        self.offsets, total = [], 0
        for k, shp in enumerate(shapes):
            a = ALIGN if aligns is None else int(aligns[k])
            total = (total + a - 1) // a * a
            self.offsets.append(total)
            total += int(np.prod(shp)) if len(shp) else 1
        total = (total + ALIGN - 1) // ALIGN * ALIGN
        self.total = total
        self.shapes = [tuple(s) for s in shapes]
        self.params = torch.zeros(total, dtype=torch.float32, device=device)
        self.grads = torch.zeros(total, dtype=torch.float32, device=device)
        self.bf16 = torch.zeros(total, dtype=torch.bfloat16, device=device) if with_bf16 else None
        self.m = None
        self.v = None

    def view(self, buf, i):
        n = int(np.prod(self.shapes[i])) if len(self.shapes[i]) else 1
        return buf[self.offsets[i]:self.offsets[i] + n].view(self.shapes[i])

    def ensure_moments(self):
        if self.m is None:
            self.m = torch.zeros_like(self.params)
            self.v = torch.zeros_like(self.params)


_FLAT_REGISTRY = {}   # storage data_ptr of a flat parameter buffer -> (FlatBuffers, {param data_ptr: index})


def register_flat(flat, params):
    _FLAT_REGISTRY[flat.params.untyped_storage().data_ptr()] = (flat, {p.data_ptr(): i for i, p in enumerate(params)})


class BertAdam(torch.optim.Optimizer):
    """Implements the BERT version of Adam with weight-decay fix (reference signature, models/optimization.py:200-201)."""

    def __init__(self, params, lr=None, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 weight_decay=0.01, max_grad_norm=1.0, **kwargs):
        if lr is None:
            raise ValueError("lr is required")
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if schedule not in ("warmup_linear", "warmup_constant", "none", None):
            raise ValueError("Invalid schedule parameter: {}".format(schedule))
        if not 0.0 <= b1 < 1.0:
            raise ValueError("Invalid b1 parameter: {} - should be in [0.0, 1.0[".format(b1))
        if not 0.0 <= b2 < 1.0:
            raise ValueError("Invalid b2 parameter: {} - should be in [0.0, 1.0[".format(b2))
        if not e >= 0.0:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(e))
        if not 0.0 <= warmup < 1.0 and not warmup == -1:
            raise ValueError("Invalid warmup: {} - should be in [0.0, 1.0[ or -1".format(warmup))
        defaults = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                        weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self._plist = [p for g in self.param_groups for p in g["params"]]
        if not self._plist or not all(p.is_cuda for p in self._plist):
            raise RuntimeError("nbest_b200.BertAdam needs CUDA parameters (no CPU fallback)")
        self._bind_flat()
        self._steps = [0] * len(self._plist)
        self._tables = None
        self._active_sig = None

    # ------------------------------------------------------------------ flat buffers
    def _bind_flat(self):
        sp = self._plist[0].untyped_storage().data_ptr()
        reg = _FLAT_REGISTRY.get(sp)
        if reg is not None and all(p.untyped_storage().data_ptr() == sp and p.data_ptr() in reg[1] for p in self._plist):
            self.flat, index = reg
            self._idx = [index[p.data_ptr()] for p in self._plist]
            return
        # foreign parameters: move them into fresh flat buffers (p.data / p.grad become views)
        flat = FlatBuffers([tuple(p.shape) for p in self._plist], self._plist[0].device, with_bf16=False)
        for i, p in enumerate(self._plist):
            flat.view(flat.params, i).copy_(p.data)
            had_grad = p.grad is not None
            if had_grad:
                flat.view(flat.grads, i).copy_(p.grad)
            p.data = flat.view(flat.params, i)
            p.grad = flat.view(flat.grads, i) if had_grad else None
        self.flat = flat
        self._idx = list(range(len(self._plist)))
        self._foreign = True
        register_flat(flat, self._plist)

    def zero_grad(self, set_to_none=False):
        """Keeps .grad tensors alive as views of the flat gradient buffer (one memset instead of 221 kernels)."""
        self.flat.grads.zero_()

    def _hyper(self):
        out = []
        for g in self.param_groups:
            for _ in g["params"]:
                out.append(g)
        return out

    def _rebuild(self, active):
        groups = self._hyper()
        spec = []
        for p, g, i, a in zip(self._plist, groups, self._idx, active):
            spec.append(dict(offset=self.flat.offsets[i], numel=p.numel(), lr=g["lr"], weight_decay=g["weight_decay"], active=a))
        self._tables = build_adam_tables(spec, self._plist[0].device)
        self._active_sig = (tuple(active), tuple((g["lr"], g["weight_decay"]) for g in groups))

    def get_lr(self):
        lr = []
        for g, s in zip(self._hyper(), self._steps):
            lr.append(g["lr"] * schedule_multiplier(s, g["t_total"], g["warmup"], g["schedule"]))
        return lr

    @property
    def state_views(self):
        """{param: {'step', 'next_m', 'next_v'}} like the reference's optimizer.state (views into the flat moments)."""
        self.flat.ensure_moments()
        return {p: dict(step=s, next_m=self.flat.view(self.flat.m, i), next_v=self.flat.view(self.flat.v, i))
                for p, s, i in zip(self._plist, self._steps, self._idx)}

    # ------------------------------------------------------------------ bucketed step (overlapped with backward)
    def set_buckets(self, segments):
        """segments: (name, start, end) element ranges of the flat buffers, e.g. trainer.model_segments(). Afterwards
        begin_bucketed_step() / step_bucket(name) / end_bucketed_step() update one bucket at a time — each as soon as its
        gradients are final — instead of everything after the last gradient (same arithmetic as step(): clipping is per
        tensor, models/optimization.py:270-271, so tensors can be stepped in any grouping)."""
        self._bucket_segments = [(n, int(s), int(e)) for n, s, e in segments]
        self._bucket_tables = None

    def _active_list(self):
        active = []
        for p, i in zip(self._plist, self._idx):
            a = p.grad is not None
            if a and p.grad.data_ptr() != self.flat.grads.data_ptr() + 4 * self.flat.offsets[i]:
                self.flat.view(self.flat.grads, i).copy_(p.grad)
                p.grad = self.flat.view(self.flat.grads, i)
            active.append(a)
        return active

    @torch.no_grad()
    def begin_bucketed_step(self):
        groups = self._hyper()
        active = self._active_list()
        sig = (tuple(active), tuple((g["lr"], g["weight_decay"]) for g in groups))
        if getattr(self, "_bucket_tables", None) is None or sig != getattr(self, "_bucket_sig", None):
            self._bucket_tables = {}
            for name, s, e in self._bucket_segments:
                spec = []
                for p, g, i, a in zip(self._plist, groups, self._idx, active):
                    off = self.flat.offsets[i]
                    spec.append(dict(offset=off, numel=p.numel(), lr=g["lr"], weight_decay=g["weight_decay"],
                                     active=a and s <= off < e))
                if any(x["active"] for x in spec):
                    self._bucket_tables[name] = build_adam_tables(spec, self._plist[0].device)
            self._bucket_sig = sig
        self.flat.ensure_moments()
        steps = {s for a, s in zip(active, self._steps) if a}
        if len(steps) > 1:
            raise RuntimeError("BertAdam: parameters became active at different steps; per-tensor step counts diverged")
        g0 = groups[0]
        self._bucket_sched = schedule_multiplier(steps.pop(), g0["t_total"], g0["warmup"], g0["schedule"]) if steps else 0.0
        self._bucket_active = active

    @torch.no_grad()
    def step_bucket(self, name):
        """Update the tensors of one bucket on the stream the caller has bound (ops.on_stream)."""
        t = self._bucket_tables.get(name)
        if t is None:
            return
        g0 = self.param_groups[0]
        ops.bertadam_step(self.flat.params, self.flat.grads, self.flat.m, self.flat.v, self.flat.bf16, t["tensors"],
                          t["n_tensors"], t["chunks"], t["n_chunks"], t["norms"], self._bucket_sched, g0["b1"], g0["b2"],
                          g0["e"], g0["max_grad_norm"])

    def end_bucketed_step(self):
        for k, a in enumerate(self._bucket_active):
            if a:
                self._steps[k] += 1

    @torch.no_grad()
    @ops.with_bound_stream
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        groups = self._hyper()
        active = []
        for p, i in zip(self._plist, self._idx):
            a = p.grad is not None
            if a and p.grad.data_ptr() != self.flat.grads.data_ptr() + 4 * self.flat.offsets[i]:
                # autograd (or user code) replaced the view: fold the foreign gradient back into the flat buffer
                self.flat.view(self.flat.grads, i).copy_(p.grad)
                p.grad = self.flat.view(self.flat.grads, i)
            active.append(a)
        sig = (tuple(active), tuple((g["lr"], g["weight_decay"]) for g in groups))
        if self._tables is None or sig != self._active_sig:
            self._rebuild(active)
        self.flat.ensure_moments()
        g0 = groups[0]
        # tensors that share a step count share one schedule multiplier; in practice all active tensors do
        by_step = {}
        for k, (a, s) in enumerate(zip(active, self._steps)):
            if a:
                by_step.setdefault(s, []).append(k)
        if len(by_step) > 1:
            raise RuntimeError("BertAdam: parameters became active at different steps; per-tensor step counts diverged")
        for s in by_step:
            sched = schedule_multiplier(s, g0["t_total"], g0["warmup"], g0["schedule"])
            ops.bertadam_step(self.flat.params, self.flat.grads, self.flat.m, self.flat.v, self.flat.bf16,
                              self._tables["tensors"], self._tables["n_tensors"], self._tables["chunks"],
                              self._tables["n_chunks"], self._tables["norms"], sched, g0["b1"], g0["b2"], g0["e"],
                              g0["max_grad_norm"])
        for k, a in enumerate(active):
            if a:
                self._steps[k] += 1
        return loss
