"""Data-parallel trainer: one process per GPU, bucketed NCCL gradient all-reduce overlapped with backward.

Replaces the reference's single-GPU picker (utils/gpu_selection.py:27-66, n_best_asr_bert.py:116-126) — the only place
the path shards. Utterances are independent through forward/backward, weights are replicated, and there
is exactly one exchange per optimizer step: the SUM of the flat fp32 gradient buffer over ranks.

Parity-critical details:
  * the BCE / CE terms are sum-reduced over the batch (n_best_asr_bert.py:572-573) -> gradients are SUMMED, no 1/R;
  * the MSE term is a mean over B*768 (n_best_asr_bert.py:574) -> each rank scales it by 1/R (mse_scale);
  * BertAdam clips per tensor (models/optimization.py:270-271) -> the norm pass runs AFTER the all-reduce;
  * the pooler never receives gradients -> it is not in any bucket.
Buckets are contiguous slices of the flat gradient buffer in the order backward finishes them (head, layer L-1 ... 0,
embeddings); each is all-reduced on a side stream as soon as its last wgrad kernel has been enqueued.
"""
import os

import torch
import torch.distributed as dist

from . import ops


class GradBucketer:
    """Contiguous gradient buckets over a flat buffer + their (optionally asynchronous) SUM all-reduce.

    `segments` is a list of (name, start, end) element ranges of the flat gradient buffer, in the order backward
    completes them. Works on CPU tensors with the gloo backend as well (used by the world_size-2 tests).
    """

    def __init__(self, flat_grads, segments, group=None, comm_stream=None):
        self.flat = flat_grads
        self.segments = list(segments)
        self.by_name = {n: (s, e) for n, s, e in self.segments}
        self.group = group
        self.comm_stream = comm_stream
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._pending = []
        self.stamp = None             # optional callable(tag, stream): the trainer's timeline recorder

    def reduce(self, name):
        """Enqueue the all-reduce of one bucket. CUDA: on the comm stream, after everything enqueued so far on the
        current stream; CPU (gloo): asynchronous work handle."""
        if self.world == 1:
            return
        if name not in self.by_name:
            return                                # a layer inside a merged bucket: its group is not complete yet
        s, e = self.by_name[name]
        buf = self.flat[s:e]
        if buf.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if self.stamp:
                    self.stamp("allreduce_start:" + name, self.comm_stream)
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                if self.stamp:
                    self.stamp("allreduce_end:" + name, self.comm_stream)
        else:
            self._pending.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        """Make the current stream (or the host, for gloo) wait for every enqueued bucket."""
        if self.world == 1:
            return
        if self.flat.is_cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        for w in self._pending:
            w.wait()
        self._pending = []


def model_segments(model, layers_per_bucket=None):
    """Bucket plan for a TOD_ASR_Transformer_STC in the order backward finishes the gradients:
    [head + pooler gap + top layers] | ... groups of `layers_per_bucket` layers ... | embeddings.
    A bucket is named after the LAST layer of its group to finish (the lowest index), which is the name the model's
    backward reports; the other layers of the group report names that are not buckets and are ignored. The pooler's
    (always zero) gradient region lies inside the first span, which keeps every bucket one contiguous slice."""
    if layers_per_bucket is None:
        layers_per_bucket = int(os.environ.get("NBEST_BUCKET_LAYERS", "2"))
    f, idx = model.flat, model._index
    L = model.spec.layers

    def start(name):
        return f.offsets[idx[name]]

    lay = lambda l: "bert_encoder.encoder.layer.%d.attention.self.query.weight" % l
    segs, hi, end = [], L, f.total
    while hi > 0:
        # the lowest layers go one per bucket: whatever is still un-reduced when the backward ends is the exposed tail of
        # the step (nothing is left to overlap with), so the last bucket before the embeddings is kept small
        lo = max(0, hi - (layers_per_bucket if hi > layers_per_bucket else 1))
        segs.append(("layer%d" % lo, start(lay(lo)), end))
        end, hi = start(lay(lo)), lo
    segs.append(("emb", 0, end))
    return segs


def init_distributed():
    """torchrun-style bootstrap (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if torch.cuda.is_available():
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, local, world


class NcclGradPool:
    """Gradient buffer in NCCL-registered memory: with ncclMemAlloc'ed, communicator-registered user buffers NCCL's NVLS
    all-reduce reads and writes the gradients in place through the NVSwitch multicast object instead of staging them
    through its own buffers with a full set of copy CTAs — fewer SMs taken from the backward the collective overlaps.

        pool = NcclGradPool(device)          # after init_distributed(), BEFORE the model is built
        with pool:
            model = TOD_ASR_Transformer_STC(...)          # its flat gradient buffer is allocated from the pool
        pool.register()                                   # ncclCommRegister of the pool's segments

    Everything is best effort: if this torch / NCCL build lacks the hooks, `ok` stays False and the model allocates as
    usual (NBEST_NCCL_POOL=0 skips the attempt)."""

    def __init__(self, device, group=None):
        self.ok, self.why, self.pool, self.backend = False, None, None, None
        if os.environ.get("NBEST_NCCL_POOL", "1") == "0":
            self.why = "NBEST_NCCL_POOL=0"
            return
        try:
            pg = group if group is not None else dist.distributed_c10d._get_default_group()
            self.backend = pg._get_backend(torch.device(device))
            self.pool = torch.cuda.MemPool(self.backend.mem_allocator)
            self.ok = True
        except Exception as e:            # no NCCL allocator hooks in this build
            self.why = repr(e)[:200]

    def __enter__(self):
        if self.ok:
            from . import optim
            optim.GRAD_POOL = self.pool
        return self

    def __exit__(self, *exc):
        from . import optim
        optim.GRAD_POOL = None
        return False

    def register(self):
        if not self.ok:
            return False
        try:
            self.backend.register_mem_pool(self.pool)
            return True
        except Exception as e:
            self.ok, self.why = False, repr(e)[:200]
            return False


class DataParallelTrainer:
    """Fused training step on one rank of a data-parallel job.

        trainer = DataParallelTrainer(model, optimizer, add_l2_loss=False)
        losses = trainer.step(ids, labels, trans_ids, seg, trans_seg)     # device tensor [mse, bce_final, bce_top, ce]

    `losses` holds this rank's local terms (sum-reduced over its own utterances); `global_losses()` all-reduces them
    for logging (once per logging interval, not per step)."""

    def __init__(self, model, optimizer, add_l2_loss=False, group=None, overlap_optimizer=None):
        self.model, self.optimizer, self.add_l2_loss = model, optimizer, add_l2_loss
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.comm_stream = torch.cuda.Stream(device=model.device) if self.world > 1 else None
        # Collective placement: a bucket whose gradients are final is not reduced at once but at the next "slot" the
        # backward announces (just before an attention-backward window). NCCL's CTAs then share the GPU with many small
        # CTAs; launched next to a persistent tcgen05 GEMM they would hold SMs that the GEMM's statically scheduled CTA
        # pairs wait for (measured: dgrad GEMMs 22 % slower under overlap). One layer per bucket (28 MB) fits a window.
        env_slots = os.environ.get("NBEST_COMM_SLOTS")
        self.comm_slots = self.world > 1 and (env_slots is None or env_slots != "0")
        # (bucket size re-measured in round 2 with PDL + dynamic GEMM scheduling, N = 2: one layer per bucket at the slots
        #  10.89 ms, two layers 10.58 ms, immediate launch 10.64 - 10.74 ms)
        segments = model_segments(model, None)
        # (One communicator. Measured at N = 4 with registered NVLS buffers: a second communicator capped at 8 / 6 CTAs for
        #  the buckets in the body of the backward makes their all-reduces 2.5 - 3x longer without making the backward any
        #  faster, and delays the exposed last buckets queued behind them: step 10.43 -> 10.88 / 10.92 ms.)
        self.bucketer = GradBucketer(model.flat.grads, segments, group, self.comm_stream)
        self.group = group
        self._pending = []
        self._emb_seen = False
        self._zeroed = 0
        # BertAdam per bucket on a side stream, each bucket as soon as its (all-reduced) gradients are final: the HBM-bound
        # update of layer l then runs under the tensor-bound backward GEMMs of the layers below it instead of after them
        # (single GPU: measured neutral — the update competes with the wgrad GEMMs for HBM — so it is on by default only
        #  for data-parallel runs, where it also takes the update off the tail behind the last all-reduce)
        if overlap_optimizer is None:
            env = os.environ.get("NBEST_OVERLAP_ADAM")
            overlap_optimizer = (self.world > 1) if env is None else env != "0"
        # (AdamW / Adam clip the GLOBAL gradient norm, which needs every gradient before the first update: no buckets)
        self.overlap_optimizer = bool(overlap_optimizer) and getattr(optimizer, "supports_buckets", False) and \
            getattr(optimizer, "flat", None) is model.flat
        # independent dropout masks per rank (the counter-based hash is keyed on seed / layer / site / element only);
        # the base seed in the checkpoint stays rank-free
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        model._seed_salt = (rank * 0x9E3779B1) & 0xFFFFFFFF
        # Under gradient all-reduce overlap NCCL's CTAs hold SMs that statically scheduled persistent GEMM pairs would wait
        # for: let the GEMMs draw their tiles dynamically (csrc/gemm_sm100.cu). Not on one GPU: there it only costs.
        if "NBEST_GEMM_DYNAMIC" not in os.environ and model.device.type == "cuda":
            from . import _lib
            _lib.context(model.device.index).set_gemm_dynamic(self.world > 1)
        self._bucket_names = {n for n, _, _ in segments}
        # Word-embedding gradient: of the V x 768 table only the rows of tokens that occurred in the batch are non-zero
        # (<= T of XLM-R's 250,002). For large tables the ranks exchange the union of touched rows as one compact
        # all-reduce instead of the dense table (768 MB fp32 for XLM-R: the un-overlappable tail of the step, since the
        # embedding gradient is the last thing backward produces). Exact: untouched rows are zero on every rank.
        env_sp = os.environ.get("NBEST_SPARSE_EMB")
        big = model.spec.vocab_size * 768 * 4 >= (256 << 20)
        self.sparse_emb = self.world > 1 and model.flat.grads.is_cuda and ((env_sp != "0") if env_sp is not None else big) \
            and (env_sp == "1" or big)
        if self.sparse_emb:
            V = model.spec.vocab_size
            self._sp_flags = torch.zeros(V, dtype=torch.int32, device=model.device)
            self._sp_rows = torch.empty(V, dtype=torch.int32, device=model.device)
            self._sp_count = torch.zeros(1, dtype=torch.int32, device=model.device)
            self._sp_buf = None
        if self.overlap_optimizer:
            optimizer.set_buckets(segments)
            self.opt_stream = torch.cuda.Stream(device=model.device)
        self.timeline = None          # start_timeline(): [(tag, CUDA event)] of one step (profiles/dp_timeline.py)

    # ------------------------------------------------------------------ per-bucket timeline (diagnostics)
    def start_timeline(self):
        """Record CUDA events at the backward's bucket / slot announcements (main stream), around every bucket's all-reduce
        (comm stream) and BertAdam update (optimizer stream) of the steps that follow; stop_timeline() returns
        [(tag, ms since the step started)]."""
        self.timeline = []
        self.bucketer.stamp = self._stamp

    def _stamp(self, tag, stream=None):
        if self.timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream if stream is not None else torch.cuda.current_stream())
            self.timeline.append((tag, ev))

    def stop_timeline(self):
        torch.cuda.synchronize()
        tl, self.timeline = self.timeline, None
        self.bucketer.stamp = None
        if not tl:
            return []
        t0 = tl[0][1]
        return [(tag, t0.elapsed_time(ev)) for tag, ev in tl]

    def _grad_ready(self, name):
        if self.timeline is not None:
            self._stamp("backward:" + name)
        if self.comm_slots:
            if name == "slot" or name == "emb":
                if name == "emb":
                    self._pending.append(name)
                    self._emb_seen = True
                pending, self._pending = self._pending, []
                for n in pending:
                    self._launch_bucket(n)
            elif name in self._bucket_names:
                if self._emb_seen:       # the lowest layer's deferred weight gradients: no later slot will come
                    self._launch_bucket(name)
                else:
                    self._pending.append(name)
            return
        if name != "slot":
            self._launch_bucket(name)

    def _reduce_emb_sparse(self):
        """Row-sparse SUM of the word-embedding gradient + dense SUM of the rest of the embedding bucket, on the comm stream."""
        m, f = self.model, self.model.flat
        V = m.spec.vocab_size
        w0 = f.offsets[m._index["bert_encoder.embeddings.word_embeddings.weight"]]
        s, e = self.bucketer.by_name["emb"]
        assert s <= w0 and w0 + V * 768 <= e
        gw = f.grads[w0:w0 + V * 768].view(V, 768)
        pk, T_act = m._last_pk, m._last_T_act
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            with ops.on_stream(self.comm_stream):
                ops.rows_mark(pk.tokens, T_act, self._sp_flags)
            dist.all_reduce(self._sp_flags, op=dist.ReduceOp.MAX, group=self.group)
            with ops.on_stream(self.comm_stream):
                ops.rows_compact(self._sp_flags, m.spec.pad_token_id, self._sp_rows, self._sp_count)
            n = int(self._sp_count.item())               # the one host read: sizes the compact exchange
            if n > 0:
                if self._sp_buf is None or self._sp_buf.shape[0] < n:
                    self._sp_buf = torch.empty((max(n, 4096) * 5 // 4, 768), dtype=torch.float32, device=m.device)
                buf = self._sp_buf[:n]
                with ops.on_stream(self.comm_stream):
                    ops.rows_move_f32(gw, self._sp_rows, n, buf, scatter=False)
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                with ops.on_stream(self.comm_stream):
                    ops.rows_move_f32(buf, self._sp_rows, n, gw, scatter=True)
            for a, b in ((s, w0), (w0 + V * 768, e)):
                if b > a:
                    dist.all_reduce(f.grads[a:b], op=dist.ReduceOp.SUM, group=self.group)
        self.last_sparse_rows = n

    def _launch_bucket(self, name):
        if name == "emb" and self.sparse_emb and getattr(self.model, "_last_pk", None) is not None:
            self._reduce_emb_sparse()
        else:
            self.bucketer.reduce(name)                   # DP: all-reduce on the comm stream after what is enqueued so far
        if not self.overlap_optimizer or name not in self._bucket_names:
            return
        ev = torch.cuda.Event()
        ev.record(self.comm_stream if self.world > 1 else torch.cuda.current_stream())
        self.opt_stream.wait_event(ev)
        with ops.on_stream(self.opt_stream):
            if self.timeline is not None:
                self._stamp("adam_start:" + name, self.opt_stream)
            self.optimizer.step_bucket(name)
            # the step's gradient zero-fill, bucket by bucket behind each update: only the last bucket's share of the
            # 438 MB memset is left after the backward (a store folded into the update kernel itself was measured: the
            # read-then-write of the same lines slows the update from 0.59 to 2.0 ms)
            s, e = self.bucketer.by_name[name]
            ops.zero_(self.model.flat.grads[s:e])
            self._zeroed += e - s
            if self.timeline is not None:
                self._stamp("adam_end:" + name, self.opt_stream)

    @ops.with_bound_stream
    def accumulate(self, ids, labels, trans_ids=None, seg=None, trans_seg=None, lens=None, trans_lens=None, n_real=None):
        """Forward + loss + backward of one micro-batch INTO the flat gradient buffer, without exchange or update: the
        first n_accum_steps - 1 micro-batches of the reference's gradient accumulation (n_best_asr_bert.py:264-266; the
        wgrad kernels accumulate with fp32 red.add, so nothing else is needed). The micro-batch that completes the
        group goes through step(), whose bucketed all-reduce then carries the accumulated sums."""
        m = self.model
        m._grad_ready_hook = None
        losses, head = m.forward_loss_backward(ids, labels, trans_ids, seg, trans_seg, add_l2_loss=self.add_l2_loss,
                                               mse_scale=1.0 / self.world, input_lens=lens, trans_input_lens=trans_lens,
                                               n_real=n_real)
        self.last_head = head
        return losses

    @ops.with_bound_stream
    def step(self, ids, labels, trans_ids=None, seg=None, trans_seg=None, lens=None, trans_lens=None, clip_norm=None,
             scheduler=None, n_real=None):
        """clip_norm: global gradient-norm clip folded into an AdamW / Adam update (n_best_asr_bert.py:268-271);
        scheduler: stepped after the update (adamw branch, :276-277); n_real: rows behind it are shape fillers
        (graph.add_fillers; model.forward_loss_backward)."""
        m = self.model
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())     # zeroed grads / previous step are visible
        if self.overlap_optimizer:
            self.opt_stream.wait_stream(torch.cuda.current_stream())      # the previous step's zero_grad is visible
            self.optimizer.begin_bucketed_step()
            m._grad_ready_hook = self._grad_ready
        else:
            m._grad_ready_hook = self._grad_ready if self.world > 1 else None
        self._pending = []
        self._emb_seen = False
        self._zeroed = 0
        if self.timeline is not None:
            self._stamp("step_start")
        try:
            losses, head = m.forward_loss_backward(ids, labels, trans_ids, seg, trans_seg, add_l2_loss=self.add_l2_loss,
                                                   mse_scale=1.0 / self.world, input_lens=lens, trans_input_lens=trans_lens,
                                                   n_real=n_real)
        finally:
            m._grad_ready_hook = None
        for n in self._pending:                          # (only if the backward never announced "emb")
            self._launch_bucket(n)
        self._pending = []
        self.bucketer.wait()
        if self.overlap_optimizer:
            torch.cuda.current_stream().wait_stream(self.opt_stream)
            self.optimizer.end_bucketed_step()
        else:
            if clip_norm is not None and clip_norm > 0:
                from .optim import clip_grad_norm_
                clip_grad_norm_(m._plist, clip_norm, optimizer=self.optimizer)
            self.optimizer.step()
        if scheduler is not None:
            scheduler.step()
        if not (self.overlap_optimizer and self._zeroed == self.model.flat.grads.numel()):
            self.optimizer.zero_grad()
        if self.timeline is not None:
            self._stamp("step_end")
        self.last_head = head
        return losses

    def global_losses(self, losses):
        if self.world > 1:
            losses = losses.clone()
            dist.all_reduce(losses, op=dist.ReduceOp.SUM, group=self.group)
        return losses
