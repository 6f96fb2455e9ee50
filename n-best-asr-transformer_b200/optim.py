"""BertAdam drop-in (reference models/optimization.py:183-302) backed by the fused multi-tensor kernel.

Same constructor and semantics as the reference class — per-tensor clip_grad_norm_(p, max_grad_norm), Adam moments
without bias correction, e = 1e-6, decoupled weight decay, `lr * WarmupLinearSchedule(step / t_total)` with the first
update at lr 0, tensors with `grad is None` skipped — but the 221 one-tensor param groups of
n_best_asr_bert.py:535-550 are updated by two kernel launches over one flat fp32 buffer instead of ~3 k launches.

`AdamW` / `Adam` are the drop-ins for the other two `--optim_choice` branches (n_best_asr_bert.py:553-569) on the same
flat buffers and kernel (nbest_adam_step), with `clip_grad_norm_` for the global clipping the reference applies to them
(:268-271), and `get_linear_schedule_with_warmup` for AdamW's scheduler (:564-568).
"""
import ctypes as C

import numpy as np
import torch

from . import ops
from . import _lib
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
                norms=torch.zeros(n + len(chunks), dtype=torch.float32, device=device))


GRAD_POOL = None      # a torch.cuda.MemPool the next FlatBuffers allocates its gradient buffer from (trainer.NcclGradPool)


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
        if GRAD_POOL is not None and torch.device(device).type == "cuda":
            # data-parallel runs: the buffer every gradient all-reduce works on comes from NCCL's own allocator and is
            # registered with the communicator (zero-copy NVLS collectives, see trainer.NcclGradPool)
            try:
                with torch.cuda.use_mem_pool(GRAD_POOL, device=torch.device(device)):
                    self.grads = torch.zeros(total, dtype=torch.float32, device=device)
            except Exception:         # best effort: NCCL's allocator refused (no multicast support, ...) -> ordinary memory
                self.grads = torch.zeros(total, dtype=torch.float32, device=device)
        else:
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
        self._init_flat()

    _mode = _lib.ADAM_BERT
    _global_clip = False
    _UNIFORM_KEYS = ("b1", "b2", "e", "max_grad_norm", "t_total", "warmup", "schedule")

    def _init_flat(self):
        self._plist = [p for g in self.param_groups for p in g["params"]]
        if not self._plist or not all(p.is_cuda for p in self._plist):
            raise RuntimeError("nbest_b200 optimizers need CUDA parameters (no CPU fallback)")
        self._bind_flat()
        self._steps = [0] * len(self._plist)
        self._tables = None
        self._active_sig = None
        self._pending_clip = None

    def _uniform(self):
        """The fused kernel takes b1 / b2 / e / max_grad_norm and the schedule once per launch: the reference honours
        them per group, so groups that disagree are rejected instead of being silently overridden."""
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            for k in self._UNIFORM_KEYS:
                if k in g0 and g.get(k) != g0.get(k):
                    raise ValueError("param groups disagree on %r (%r vs %r): the fused step needs one value" % (k, g0.get(k), g.get(k)))
        return g0

    # ------------------------------------------------------------------ torch.optim.Optimizer state plumbing
    def state_dict(self):
        """Standard layout (`state` / `param_groups`) with the flat moments inside, so that the usual
        torch.save(optimizer.state_dict()) round-trips the Adam moments and step counts (checkpoint.py stores the same)."""
        self.flat.ensure_moments()
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        return dict(state=dict(flat_m=self.flat.m.detach().clone(), flat_v=self.flat.v.detach().clone(),
                               steps=list(self._steps), offsets=[self.flat.offsets[i] for i in self._idx],
                               numels=[p.numel() for p in self._plist]),
                    param_groups=groups)

    def load_state_dict(self, sd):
        st = sd["state"]
        if list(st["numels"]) != [p.numel() for p in self._plist] or st["flat_m"].numel() != self.flat.total:
            raise ValueError("optimizer state does not match this parameter layout")
        self.flat.ensure_moments()
        self.flat.m.copy_(st["flat_m"])
        self.flat.v.copy_(st["flat_v"])
        self._steps = [int(x) for x in st["steps"]]
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in sg.items() if k != "params"})
        self._tables = None
        if hasattr(self, "_bucket_tables"):
            self._bucket_tables = None

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
        ops.zero_(self.flat.grads)

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

    def get_lr(self):
        lr = []
        for g, s in zip(self._hyper(), self._steps):
            lr.append(g["lr"] * schedule_multiplier(s, g.get("t_total", -1), g.get("warmup", -1), g.get("schedule", "none")))
        return lr

    @property
    def supports_buckets(self):
        return not self._global_clip

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
        g0 = self._uniform()
        self._bucket_kw = self._kernel_args(g0, steps.pop()) if steps else None
        self._bucket_active = active

    @torch.no_grad()
    def step_bucket(self, name):
        """Update the tensors of one bucket on the stream the caller has bound (ops.on_stream)."""
        t = self._bucket_tables.get(name)
        if t is None or self._bucket_kw is None:
            return
        if self._global_clip:
            raise RuntimeError("global gradient clipping needs every gradient: per-bucket steps are BertAdam-only")
        ops.bertadam_step(self.flat.params, self.flat.grads, self.flat.m, self.flat.v, self.flat.bf16, t["tensors"],
                          t["n_tensors"], t["chunks"], t["n_chunks"], t["norms"], **self._bucket_kw)

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
        lr_scale = self._lr_scale(groups, active)
        self.flat.ensure_moments()
        g0 = self._uniform()
        # tensors that share a step count share one schedule multiplier; in practice all active tensors do
        by_step = {}
        for k, (a, s) in enumerate(zip(active, self._steps)):
            if a:
                by_step.setdefault(s, []).append(k)
        if len(by_step) > 1:
            raise RuntimeError("%s: parameters became active at different steps; per-tensor step counts diverged" % type(self).__name__)
        for s in by_step:
            kw = self._kernel_args(g0, s)
            kw["sched"] *= lr_scale
            ops.bertadam_step(self.flat.params, self.flat.grads, self.flat.m, self.flat.v, self.flat.bf16,
                              self._tables["tensors"], self._tables["n_tensors"], self._tables["chunks"],
                              self._tables["n_chunks"], self._tables["norms"], **kw)
        for k, a in enumerate(active):
            if a:
                self._steps[k] += 1
        self._pending_clip = None
        return loss

    def _lr_scale(self, groups, active):
        """(Re)build the device tensor table when the active set / weight decays change. Learning rates that all moved
        by ONE common factor since the table was built (what a torch LambdaLR scheduler does every step,
        n_best_asr_bert.py:564-568) are passed to the kernel as a multiplier instead of re-uploading the table."""
        wds = tuple(g["weight_decay"] for g in groups)
        lrs = [float(g["lr"]) for g in groups]
        scale = None
        if self._tables is not None and self._active_sig == (tuple(active), wds):
            base = self._table_lrs
            ratios = {round(l / b, 12) if b != 0 else (0.0 if l == 0 else None) for l, b in zip(lrs, base)}
            if len(ratios) == 1 and None not in ratios:
                scale = ratios.pop()
        if scale is None:
            self._rebuild(active)
            self._active_sig = (tuple(active), wds)
            self._table_lrs = lrs
            scale = 1.0
        return scale

    def _kernel_args(self, g0, step):
        return dict(sched=schedule_multiplier(step, g0["t_total"], g0["warmup"], g0["schedule"]), b1=g0["b1"], b2=g0["b2"],
                    eps=g0["e"], max_grad_norm=g0["max_grad_norm"], mode=self._mode, global_clip=self._global_clip, step=step + 1)


class AdamW(BertAdam):
    """transformers(2.3.0).optimization.AdamW as n_best_asr_bert.py:563 builds it — `AdamW(grouped_parameters, lr=opt.lr,
    correct_bias=False)`: p -= lr m/(sqrt(v)+eps), then decoupled decay p -= lr wd p — on the fused kernel. The
    reference clips the GLOBAL gradient norm before the step (`clip_grad_norm_(params, opt.max_norm)`, :268-271); call
    `nbest_b200.optim.clip_grad_norm_(params, max_norm, optimizer=this)` in its place and the clip is folded into the
    step's own norm pass (no extra pass over the gradients), or pass max_grad_norm= to clip on every step."""

    _mode = _lib.ADAM_HF_ADAMW
    _global_clip = True
    _UNIFORM_KEYS = ("betas", "eps", "correct_bias")

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=False, max_grad_norm=0.0):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {} - should be in [0.0, 1.0[".format(betas))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        if correct_bias:
            raise ValueError("correct_bias=True is not built: the reference constructs AdamW(correct_bias=False) "
                             "(n_best_asr_bert.py:563); use nbest_b200.optim.Adam for bias-corrected moments")
        torch.optim.Optimizer.__init__(self, params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                                          correct_bias=False))
        self.max_grad_norm = float(max_grad_norm)
        self._init_flat()

    def _kernel_args(self, g0, step):
        clip = self._pending_clip if self._pending_clip is not None else self.max_grad_norm
        return dict(sched=1.0, b1=g0["betas"][0], b2=g0["betas"][1], eps=g0["eps"], max_grad_norm=clip, mode=self._mode,
                    global_clip=True, step=step + 1)


class Adam(AdamW):
    """torch.optim.Adam as n_best_asr_bert.py:554 builds it (`betas=(0.9, 0.999), eps=1e-8, weight_decay=opt.l2`):
    L2-coupled decay, bias-corrected moments."""

    _mode = _lib.ADAM_TORCH
    _UNIFORM_KEYS = ("betas", "eps")

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        torch.optim.Optimizer.__init__(self, params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = float(max_grad_norm)
        self._init_flat()


def clip_grad_norm_(parameters, max_norm, optimizer=None):
    """Stand-in for `torch.nn.utils.clip_grad_norm_(params, opt.max_norm)` at n_best_asr_bert.py:268-271 when the
    optimizer is one of this module's AdamW / Adam: nothing is launched here — the optimizer's next step() computes the
    global norm in its own first pass and scales the gradients while it reads them (same arithmetic: coefficient
    min(1, max_norm / (||g||_2 + 1e-6)) over every parameter with a gradient). Without `optimizer` it falls back to
    torch's implementation on the .grad views."""
    if optimizer is None or not getattr(optimizer, "_global_clip", False):
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm)
    optimizer._pending_clip = float(max_norm)
    return None


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, last_epoch=-1):
    """transformers.get_linear_schedule_with_warmup (n_best_asr_bert.py:564-568): LambdaLR with
    lambda(s) = s / max(1, warmup) for s < warmup, else max(0, (total - s) / max(1, total - warmup))."""
    def lr_lambda(step):
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        return max(0.0, float(num_training_steps - step) / float(max(1, num_training_steps - num_warmup_steps)))
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda, last_epoch)
