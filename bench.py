#!/usr/bin/env python
"""Headline benchmark: train utterances/s (fwd + bwd + BertAdam step), BERT-base n-best STC, on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

BASELINE.json configs -> flags (default = configs[1], the configuration the metric is quoted on):
    configs[1]  (default)                                         BERT-base 5-best, max_len 128, B 256/GPU, training
    configs[2]  --model xlmr                                      XLM-R-base (250 k vocab), data-parallel training
    configs[3]  --l2                                              + transcript stream with gradients and the MSE term
    configs[4]  --mode infer --hyps 10 --max-len 512 --batch 512  10-best inference (forward + decode), N replicas

Workload (BASELINE.json configs[1]): BERT-base-uncased, 5-best hypotheses [SEP]-joined, max_len 128, batch 256 per GPU,
synthetic DSTC2-shaped token ids (nbest_b200.synth), random-init weights, act-slot + value heads, both encoder streams as
the reference runs them (transcript stream forward-only without --add_l2_loss), dropout on (0.1 / 0.1 / 0.3), BertAdam
with the reference's per-tensor groups. One "step" = one optimizer step over one batch. Prints ONE JSON line (rank 0).

  value      whole-job utterances/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same through the public trainer call with HOST (pinned) input buffers: H2D copies of the step's inputs
             and a D2H read of the loss terms inside the timed region
  roofline   dominant kernel = the tcgen05 GEMM (all instances of one step): algorithmic FLOPs / CUDA-event time of those
             launches, against the measured sustained bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference path timed on this box's host cores on a bounded sample (rank 0, N=1)
  gpu_eager_baseline  (informational) the unmodified reference on the SAME B200 through stock PyTorch, fp32 as written and
             with the HF encoder under autocast(bf16) (SURVEY §8(d) secondary baseline)
--impl reference times THE REFERENCE ITSELF — the unmodified modules staged under oracle/_ref/ by oracle/vendor_ref.py
(git-ignored, travels with the snapshot) — on the box's host cores, B = 32 per step, with the
fwd / bwd / optimizer split. Only if nothing is staged does it fall back to the oracle port (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train utterances/s (fwd+bwd+step) BERT-base n-best STC"
UNIT = "utterances/s"
GEMM_FLOPS_PER_TOKEN_FWD = 169_869_312          # 2*(4*768^2 + 2*768*3072)*12   
ATTN_FLOPS_PER_L2_FWD = 36_864                  # 4*768*12 per L^2


def load_hierarchy():
    with open(os.path.join(ROOT, "tests", "golden", "dstc2_hierarchy.json")) as f:
        return json.load(f)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_sustained=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf_sustained=1400.0, tf_burst=1590.0, src="fallback")


def gemm_traffic():
    """DRAM bytes (read + write) per GEMM launch, averaged over the 144 GEMM launches of one B = 256 training step, from
    the committed ncu capture (profiles/r1_gemm_dram_traffic.txt); None when the capture is absent or the workload differs."""
    p = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ CPU port (oracle)
def cpu_port_step_time(n_utt, steps, warmup, seed=999, threads=None):
    """Times the oracle port of the reference training step (both streams, dropout on, fp32, BertAdam) on host cores."""
    from oracle import stc_oracle as O
    from nbest_b200.synth import synth_batch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    hj = load_hierarchy()
    hier = O.Hierarchy({int(k): v for k, v in hj["top2bottom"].items()}, hj["none_bottoms"])
    cfg = O.EncoderConfig.bert_base()
    params = O.init_params(cfg, hier, seed=seed, style="hf")
    state = {}
    hp = dict(lr=3e-5, bert_lr=3e-5, warmup=0.1, t_total=2300, add_l2_loss=False)

    def drop(t, site):
        return torch.nn.functional.dropout(t, 0.3 if site == "head" else 0.1, True)

    times = []
    for i in range(warmup + steps):
        b = synth_batch("bert", cfg.vocab_size, hier, n_utt, 5, 128, seed + i)
        t0 = time.perf_counter()
        O.train_step(params, cfg, hier, b, state, hp, drop)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return float(np.median(times)), float(np.sum(times)), threads


def workload_name(args):
    enc = "XLM-RoBERTa-base (250k vocab)" if args.model == "xlmr" else "BERT-base-uncased"
    if args.mode == "infer":
        return "%s %d-hypothesis n-best, max_len %d, batch %d/GPU inference (forward + decode), bf16 packed varlen" % (
            enc, args.hyps, args.max_len, args.batch)
    return "%s n-best STC bf16 packed varlen, batch %d/GPU, %d hyps, max_len %d%s%s" % (
        enc, args.batch, args.hyps, args.max_len, ", dense" if args.dense else "", ", add_l2_loss" if args.l2 else "")


def metric_name(args):
    enc = "XLM-R-base" if args.model == "xlmr" else "BERT-base"
    if args.mode == "infer":
        return "inference utterances/s (fwd+decode) %s n-best STC" % enc
    return "train utterances/s (fwd+bwd+step) %s n-best STC" % enc


def reference_cpu(args, n_utt, steps, warmup):
    """(utt/s, ms/step, cpu_baseline dict): the staged reference on the host cores, else the oracle port."""
    from oracle import ref_loader
    kind = "xlm-roberta" if args.model == "xlmr" else "bert"
    if ref_loader.available():
        from oracle import ref_bench
        if args.mode == "infer":
            r = ref_bench.time_infer("cpu", n_utt, steps, warmup, kind, args.hyps, args.max_len)
            split = None
        else:
            r = ref_bench.time_train("cpu", n_utt, steps, warmup, kind, args.l2, args.hyps, args.max_len)
            split = dict(fwd_ms=r["fwd_s"] * 1e3, bwd_ms=r["bwd_s"] * 1e3, optimizer_ms=r["opt_s"] * 1e3)
        v = n_utt / r["median_s"]
        cb = dict(value=v, unit=UNIT, cores=r["threads"], kind="reference",
                  sample="the unmodified reference (oracle/_ref: make_model + cal_total_loss + backward + BertAdam, fp32, dropout "
                         "on, both streams) on %d-utterance batches of the same generator, %d warm-up + %d timed steps, %.1f s" % (
                             n_utt, warmup, steps, r["total_s"]), ms_per_step=r["median_s"] * 1e3)
        if split:
            cb["split"] = split
        return v, r["median_s"] * 1e3, cb
    med, total, threads = cpu_port_step_time(n_utt, steps, warmup)
    v = n_utt / med
    return v, med * 1e3, dict(value=v, unit=UNIT, cores=threads, kind="port", ms_per_step=med * 1e3,
                              sample="oracle port (reference not staged), %d-utterance batches, %d timed steps" % (n_utt, steps))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_utt = 32 if args.mode == "train" else 16                     # batch 32 (configs[0])
    v, ms, cb = reference_cpu(args, n_utt, args.steps, max(1, min(args.warmup, 2)))
    line = dict(impl="reference", metric=metric_name(args), value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=dict(workload=workload_name(args), global_batch=args.batch, parallelism="cpu",
                            sample="each timed step = a %d-utterance batch of that workload (same generator) on the host cores" % n_utt,
                            device="cpu"),
                cpu_baseline=cb, e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def gpu_eager_baseline(args, dev):
    """The unmodified reference on the same GPU through stock PyTorch kernels (informational)."""
    from oracle import ref_loader
    if not ref_loader.available() or args.mode != "train":
        return None
    from oracle import ref_bench
    kind = "xlm-roberta" if args.model == "xlmr" else "bert"
    out = {}
    for name, ac in (("fp32_as_written", False), ("encoder_autocast_bf16", True)):
        try:
            r = ref_bench.time_train(str(dev), args.batch, 3, 2, kind, args.l2, args.hyps, args.max_len, autocast_encoder=ac)
            out[name] = dict(value=args.batch / r["median_s"], unit=UNIT, ms_per_step=r["median_s"] * 1e3,
                             fwd_ms=r["fwd_s"] * 1e3, bwd_ms=r["bwd_s"] * 1e3, optimizer_ms=r["opt_s"] * 1e3)
        except Exception as e:      # informational leg: never fails the bench
            out[name] = dict(error=repr(e)[:200])
        torch.cuda.empty_cache()
    out["note"] = "reference code + HF encoder + torch %s kernels on this GPU, batch %d, padded, both streams, 221-group BertAdam" % (
        torch.__version__, args.batch)
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().split("\n") if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm),
                    power_w_max=float(max(power)))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU per step")
    ap.add_argument("--hyps", type=int, default=5)
    ap.add_argument("--max-len", type=int, default=128)
    ap.add_argument("--dense", action="store_true", help="every sequence exactly max_len tokens (roofline worst case)")
    ap.add_argument("--l2", action="store_true", help="--add_l2_loss: transcript stream with gradients + MSE term")
    ap.add_argument("--model", default="bert", choices=["bert", "xlmr"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the training step as a CUDA graph (nbest_b200.graph; one GPU, BERT packing, training). auto: "
                         "time both the eager and the graphed step and report the faster one, naming it in config.step_mode")
    ap.add_argument("--bucket", default="3,256", help="shape buckets of the graphed step: fillers per stream, token multiple")
    ap.add_argument("--skip-transcript", action="store_true",
                    help="do not run the transcript stream at all (without --add_l2_loss the reference computes it forward-only "
                         "and never uses the result; default = run it, as the reference does)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    from nbest_b200 import _lib, ops
    from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC
    from nbest_b200.optim import BertAdam
    from nbest_b200.synth import synth_batch
    from nbest_b200.trainer import DataParallelTrainer, init_distributed

    # NCCL prints its version banner to stdout from C; stdout must carry exactly one JSON line, so everything up to
    # the final print goes to stderr.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU port)")
    rank, local, world = init_distributed()
    dev = torch.device("cuda", local)
    hj = load_hierarchy()
    t2b = {int(k): v for k, v in hj["top2bottom"].items()}
    drop = 0.0 if args.no_dropout else None
    mk = EncoderSpec.xlmr_base if args.model == "xlmr" else EncoderSpec.bert_base
    kind = "xlm-roberta" if args.model == "xlmr" else "bert"
    spec = mk() if drop is None else mk(hidden_dropout=0.0, attn_dropout=0.0)
    infer = args.mode == "infer"
    # data-parallel: the flat gradient buffer lives in NCCL-registered memory (zero-copy NVLS all-reduce)
    from nbest_b200.trainer import NcclGradPool
    gpool = NcclGradPool(dev) if (world > 1 and not infer) else None
    if gpool is not None:
        gpool.__enter__()
    model = TOD_ASR_Transformer_STC(spec=spec, top2bottom=t2b, dropout=0.3 if drop is None else 0.0, device=dev,
                                    none_bottoms=hj["none_bottoms"], seed=999)
    if gpool is not None:
        gpool.__exit__()
        gpool.register()
        print("rank %d: gradient buffer in NCCL-registered memory: %s %s" % (rank, gpool.ok, gpool.why or ""), file=sys.stderr)
    model.train(not infer)
    groups = []
    for n, p in model.named_parameters():                          # reference n_best_asr_bert.py:535-550
        no_decay = any(nd in n for nd in ("bias", "LayerNorm.bias", "LayerNorm.weight"))
        groups.append(dict(params=p, weight_decay=0.0 if no_decay else 0.01, lr=3e-5))
    optim = BertAdam(groups, lr=3e-5, warmup=0.1, t_total=2300)
    trainer = DataParallelTrainer(model, optim, add_l2_loss=args.l2)
    ctx = _lib.context(local)

    # ---- synthetic batches: pinned host copies (e2e) + device-resident copies (value)
    NB = 4
    host, devb, stats, len_max, tlen_max = [], [], [], [], []
    keys = ("ids", "seg", "trans_ids", "trans_seg", "labels")
    for i in range(NB):
        b = synth_batch(kind, spec.vocab_size, model.hier, args.batch, args.hyps, args.max_len, seed=999 + 1000 * rank + i,
                        dense=args.dense)
        host.append({k: b[k].pin_memory() for k in keys} | dict(lens=b["lens"], trans_lens=b["trans_lens"]))
        devb.append({k: b[k].to(dev) for k in keys} | dict(lens=b["lens"], trans_lens=b["trans_lens"]))
        L = np.array(b["lens"], dtype=np.float64)
        Lt = np.array(b["trans_lens"], dtype=np.float64)
        stats.append((L.sum(), (L ** 2).sum(), Lt.sum(), (Lt ** 2).sum()))
        len_max.append(int(b["ids"].shape[1]))
        tlen_max.append(int(b["trans_ids"].shape[1]))
    h2d_bytes = int(np.mean([sum(h[k].numel() * h[k].element_size() for k in keys) for h in host]))

    skip_t = args.skip_transcript and not args.l2
    if infer:
        from nbest_b200.epoch import EpochMetrics
        keys = ("ids", "seg", "labels")
        h2d_bytes = int(np.mean([sum(h[k].numel() * h[k].element_size() for k in keys) for h in host]))
        metrics = EpochMetrics(dev)
        dec_host = torch.empty((args.batch, model.hier.n_bottom), dtype=torch.uint8).pin_memory()

    def infer_dev(i):
        # eval_epoch's hot path (n_best_asr_bert.py:316-350) without the loss: forward, decode bitmap, device-side counters
        b = devb[i % NB]
        head = model.infer(b["ids"], b["seg"], b["lens"])
        metrics.update(head.decode, b["labels"])
        return head.decode

    def infer_host(i):
        h = host[i % NB]
        d = {k: h[k].to(dev, non_blocking=True) for k in keys}
        head = model.infer(d["ids"], d["seg"], h["lens"])
        metrics.update(head.decode, d["labels"])
        dec_host.copy_(head.decode, non_blocking=True)              # D2H of the step's result: the prediction bitmap
        torch.cuda.current_stream().synchronize()
        return dec_host

    def step_dev(i):
        if infer:
            return infer_dev(i)
        b = devb[i % NB]
        if skip_t:
            return trainer.step(b["ids"], b["labels"], None, b["seg"], None, b["lens"], None)
        return trainer.step(b["ids"], b["labels"], b["trans_ids"], b["seg"], b["trans_seg"], b["lens"], b["trans_lens"])

    def step_host(i):
        if infer:
            return infer_host(i)
        h = host[i % NB]
        d = {k: h[k].to(dev, non_blocking=True) for k in keys}
        if skip_t:
            losses = trainer.step(d["ids"], d["labels"], None, d["seg"], None, h["lens"], None)
        else:
            losses = trainer.step(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], h["lens"], h["trans_lens"])
        return losses.cpu()                                         # D2H read of the step's loss terms (synchronises)

    # ---- graphed step (one GPU): the same batches with filler sequences appended by the data pipeline ahead of the timed
    # region (like collation / pinning), so that token counts fall on shape buckets and every batch replays a captured graph
    use_graph = args.graph != "off" and world == 1 and args.model == "bert" and not infer
    gt = None
    if use_graph:
        from nbest_b200.graph import GraphedTrainer, add_fillers
        n_fill, mult = (int(x) for x in args.bucket.split(","))
        gt = GraphedTrainer(trainer)
        hostg, devg = [], []
        for h in host:
            ids, seg, lens = add_fillers(h["ids"], h["seg"], h["lens"], n_fill, mult, args.max_len)
            g = dict(ids=ids, seg=seg, lens=lens, labels=h["labels"], trans_ids=None, trans_seg=None, trans_lens=None)
            if not skip_t:
                g["trans_ids"], g["trans_seg"], g["trans_lens"] = add_fillers(h["trans_ids"], h["trans_seg"], h["trans_lens"],
                                                                              n_fill, mult, args.max_len)
            hostg.append({k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in g.items()})
            devg.append({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in g.items()})
        h2d_bytes_graph = int(np.mean([sum(h[k].numel() * h[k].element_size() for k in keys if h[k] is not None) for h in hostg]))

    def step_graph_dev(i):
        b = devg[i % NB]
        return gt.step(b["ids"], b["labels"], b["trans_ids"], b["seg"], b["trans_seg"], b["lens"], b["trans_lens"], n_real=args.batch)

    def step_graph_host(i):
        h = hostg[i % NB]       # pinned host tensors: the copies into the graph's static inputs are the step's H2D traffic
        losses = gt.step(h["ids"], h["labels"], h["trans_ids"], h["seg"], h["trans_seg"], h["lens"], h["trans_lens"],
                         n_real=args.batch)
        return losses.cpu()                                         # D2H read of the step's loss terms (synchronises)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches()
        e0.record()
        for i in range(steps):
            out = fn(i)
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ctx.launches() - l0, out

    # bring-up before the W warm-up steps: every distinct batch shape once (the caching allocator grows on first use and
    # cudaMalloc synchronises), and the NCCL communicator (NVLS buffers, channel setup) when world > 1
    for i in range(NB + (4 if world > 1 else 0)):
        step_dev(i)
    for i in range(args.warmup):
        step_dev(i)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, last = timed(step_dev, args.steps)
    for i in range(2):
        step_host(i)
    ms_e2e, _, last_losses = timed(step_host, args.steps)
    modes = dict(eager=dict(ms_per_step=ms / args.steps, e2e_ms_per_step=ms_e2e / args.steps, gpu_launches=int(launches)))
    step_mode = "eager (one Python-issued launch per kernel)"
    if gt is not None:
        for i in range(NB + 1 + args.warmup):                        # first call eager, then one capture per batch shape
            step_graph_dev(i)
        k0 = gt.kernels_total
        ms_g, _, _ = timed(step_graph_dev, args.steps)
        launches_g = gt.kernels_total - k0
        for i in range(2):
            step_graph_host(i)
        ms_e2e_g, _, last_g = timed(step_graph_host, args.steps)
        modes["graph"] = dict(ms_per_step=ms_g / args.steps, e2e_ms_per_step=ms_e2e_g / args.steps, gpu_launches=int(launches_g),
                              graphs_captured=gt.captures, replays=gt.replays, eager_fallbacks=gt.eager_steps,
                              capture_error=gt.capture_error, h2d_bytes_per_step=h2d_bytes_graph,
                              shape_buckets="%d filler sequences per stream, token counts rounded up to multiples of %d "
                                            "(appended by the data pipeline ahead of the timed region)" % (n_fill, mult))
        if gt.captures > 0 and gt.capture_error is None and (args.graph == "on" or ms_e2e_g < ms_e2e):
            ms, ms_e2e, launches, last_losses, h2d_bytes = ms_g, ms_e2e_g, launches_g, last_g, h2d_bytes_graph
            step_mode = "cuda-graph replay of the whole step (nbest_b200.graph.GraphedTrainer)"
    clocks = sampler.stop() if sampler else None
    if not infer and not bool(torch.isfinite(last_losses).all()):
        raise SystemExit("non-finite loss in the benchmark step")
    d2h_bytes = int(dec_host.numel()) if infer else 16

    # ---- per-kernel roofline leg: two instrumented steps (CUDA events around every launch of ours)
    # (the per-bucket BertAdam normally runs on a side stream under the backward; the instrumented steps serialise it on
    #  the main stream so that the CUDA events around each launch measure that launch)
    overlap = trainer.overlap_optimizer
    trainer.overlap_optimizer = False
    ops.profile_start()
    for i in range(2):
        step_dev(i)
    rec = ops.profile_stop()
    trainer.overlap_optimizer = overlap
    agg = {}
    for name, t, fl, by, meta in rec:
        a = agg.setdefault(name, [0.0, 0.0, 0.0, 0])
        a[0] += t / 2
        a[1] += fl / 2
        a[2] += by / 2
        a[3] += 1
    n_params = sum(p.numel() for p in model.parameters() if p.grad is not None)
    if "bertadam_step" in agg:
        agg["bertadam_step"][2] = 32.0 * n_params

    if rank == 0:
        pk = peaks()
        utt = world * args.batch * args.steps
        value = utt / (ms / 1e3)
        T, L2, Tt, Lt2 = np.mean([s[0] for s in stats]), np.mean([s[1] for s in stats]), np.mean([s[2] for s in stats]), \
            np.mean([s[3] for s in stats])
        if kind != "bert":          # the reference leaves <pad> attendable for XLM-R: every row is real work at the batch maximum
            S_, St_ = np.asarray(len_max, dtype=np.float64), np.asarray(tlen_max, dtype=np.float64)
            T, L2, Tt, Lt2 = args.batch * S_.mean(), args.batch * (S_ ** 2).mean(), args.batch * St_.mean(), args.batch * (St_ ** 2).mean()
        if infer:
            alg = alg_all = GEMM_FLOPS_PER_TOKEN_FWD * T + ATTN_FLOPS_PER_L2_FWD * L2
            Tt = 0.0
        else:
            alg = 3.0 * (GEMM_FLOPS_PER_TOKEN_FWD * T + ATTN_FLOPS_PER_L2_FWD * L2)                  # BASELINE.md formula (ASR stream)
            alg_all = alg + (3.0 if args.l2 else (0.0 if skip_t else 1.0)) * (GEMM_FLOPS_PER_TOKEN_FWD * Tt + ATTN_FLOPS_PER_L2_FWD * Lt2)
        step_s = ms / 1e3 / args.steps
        gemm_ms = sum(agg[k][0] for k in agg if k.startswith("gemm"))
        gemm_fl = sum(agg[k][1] for k in agg if k.startswith("gemm"))
        achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        kernels = {k: dict(ms_per_step=round(v[0], 4), launches_per_step=v[3] // 2,
                           tflops=round(v[1] / (v[0] * 1e-3) / 1e12, 1) if v[1] else None,
                           gbs=round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[2] else None) for k, v in sorted(agg.items())}
        line = dict(
            metric=metric_name(args), value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
            config=dict(workload=workload_name(args),
                global_batch=world * args.batch, tokens_per_step_asr=float(T), tokens_per_step_transcript=float(Tt),
                parallelism=("replicas%d" if infer else "dp%d") % world,
                nccl_registered_grads=bool(gpool is not None and gpool.ok),
                dropout="off" if (args.no_dropout or infer) else "0.1/0.1/0.3",
                streams="asr forward + decode (eval mode)" if infer else "asr fwd+bwd, transcript %s" % (
                    "fwd+bwd" if args.l2 else ("skipped (--skip-transcript)" if skip_t else "fwd only (as the reference)")),
                step_mode=step_mode,
                l2_flush="per-step working set (activations + fp32 weights%s, > 1 GB) exceeds the 126 MB L2" % (
                    "" if infer else " + Adam state")),
            clocks=clocks,
            e2e=dict(value=utt / (ms_e2e / 1e3), unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=d2h_bytes,
                     ms_per_step=ms_e2e / args.steps),
            gpu_launches=int(launches),
            roofline=dict(bound="tensor", kernel="gemm_kernel (tcgen05, all instances of one step)", achieved=achieved,
                          peak=pk["tf_sustained"], unit="TFLOP/s", frac=achieved / pk["tf_sustained"], traffic=gemm_traffic() if (args.batch == 256 and args.hyps == 5 and not args.dense and not args.l2 and not infer and
                                                     args.model == "bert") else None,
                          peak_source=pk["src"] + " bf16_tflops_sustained", gemm_ms_per_step=gemm_ms,
                          gemm_share_of_step=gemm_ms / (step_s * 1e3)),
            modes=modes,
            tc_util=dict(asr_stream_formula=alg / (step_s * pk["tf_sustained"] * 1e12),
                         all_executed_streams=alg_all / (step_s * pk["tf_sustained"] * 1e12)),
            kernels=kernels)
        if world == 1 and not args.no_eager_baseline:
            del trainer, optim, devb
            torch.cuda.empty_cache()
            eb = gpu_eager_baseline(args, dev)
            if eb is not None:
                line["gpu_eager_baseline"] = eb
        if world == 1 and not args.no_cpu_baseline:
            _, _, line["cpu_baseline"] = reference_cpu(args, 32 if not infer else 16, 3 if not infer else 2, 1)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
