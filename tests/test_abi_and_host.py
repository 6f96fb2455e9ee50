"""CPU: the C-ABI library loads and exports every symbol include/nbest_sm100.h declares (no compute calls), and the
host-side logic (optimizer tables, schedules, flat layout, bucket plan, synthetic generator, no-fallback behaviour)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _declared_symbols():
    header = open(os.path.join(ROOT, "include", "nbest_sm100.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(nbest_[a-z0-9_]+)\s*\(", header)))


def test_library_exports_every_declared_symbol():
    from nbest_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python n-best-asr-transformer_b200/build.py` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libnbest_sm100.so does not export %s" % name
    assert sorted(_lib.exported_symbols()) == declared, "ctypes signature table and header disagree"
    lib.nbest_abi_version.restype = ctypes.c_int
    assert lib.nbest_abi_version() == 3


def test_header_cites_reference_for_every_entry_point():
    header = open(os.path.join(ROOT, "include", "nbest_sm100.h")).read()
    for frag in ("models/model.py", "utils/bert_xlnet_inputs.py", "models/optimization.py", "hierarchical_classifier.py",
                 "n_best_asr_bert.py", "modeling_bert.py", "utils/gpu_selection.py"):
        assert frag in header, frag


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    """Without a GPU the product refuses to run instead of silently computing somewhere else."""
    from nbest_b200 import _lib, ops
    from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC
    with pytest.raises(_lib.NbestError):
        _lib.Context(0)
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):
        TOD_ASR_Transformer_STC(spec=EncoderSpec.bert_base(layers=1), top2bottom={0: [0], 1: [1, 2]}, dropout=0.0, device="cpu")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "n-best-asr-transformer_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|oracle\.", src, flags=re.M), \
                fn + " imports the oracle"


def test_schedule_matches_reference_formula():
    from nbest_b200.optim import schedule_multiplier
    from oracle import stc_oracle as O
    for t_total, warm in ((2300, 0.1), (8, 0.1), (50, 0.25)):
        for step in range(0, t_total + 5):
            assert schedule_multiplier(step, t_total, warm) == O.warmup_linear(step, t_total, warm)
    assert schedule_multiplier(0, 2300, 0.1) == 0.0                       # the first update has lr 0
    assert schedule_multiplier(7, -1, 0.1) == 1.0


def test_flat_layout_alignment_and_fused_views():
    from nbest_b200.optim import FlatBuffers
    shapes = [(30, 768), (75, 768), (30,), (75,), (27,), (768,)]
    fb = FlatBuffers(shapes, "cpu", with_bf16=False, aligns=[64, 64, 64, 1, 1, 64])
    assert fb.offsets[0] == 0 and fb.offsets[1] == 30 * 768                # weights contiguous: fused [105,768] view
    assert fb.offsets[3] == fb.offsets[2] + 30 and fb.offsets[4] == fb.offsets[3] + 75     # packed biases: [132] vector
    assert fb.offsets[5] % 64 == 0 and fb.total % 64 == 0
    fb.view(fb.params, 3).fill_(2.0)
    assert float(fb.params[fb.offsets[2]:fb.offsets[2] + 132].sum()) == 150.0


def test_adam_tables_cover_active_tensors_exactly():
    from nbest_b200.optim import build_adam_tables
    spec = [dict(offset=0, numel=100000, lr=1e-3, weight_decay=0.01, active=True),
            dict(offset=100032, numel=4096, lr=1e-3, weight_decay=0.0, active=False),
            dict(offset=104128, numel=30, lr=2e-3, weight_decay=0.0, active=True),
            dict(offset=104158, numel=75, lr=2e-3, weight_decay=0.0, active=True)]
    t = build_adam_tables(spec, "cpu", chunk=16384)
    ch = t["chunks"].numpy()
    assert set(ch[:, 0].tolist()) == {0, 2, 3}
    for i, s in enumerate(spec):
        rows = ch[ch[:, 0] == i]
        if not s["active"]:
            assert len(rows) == 0
            continue
        assert rows[:, 2].sum() == s["numel"] and rows[0, 1] == 0
        assert (rows[1:, 1] == np.cumsum(rows[:-1, 2])).all()
    assert t["tensors"].numel() == 4 * ctypes.sizeof(__import__("nbest_b200._lib", fromlist=["AdamTensor"]).AdamTensor)


def test_synthetic_generator_matches_dstc2_shape_statistics():
    from nbest_b200.synth import synth_batch
    from oracle import stc_oracle as O
    hj = json.load(open(os.path.join(GOLD, "dstc2_hierarchy.json")))
    hier = O.Hierarchy({int(k): v for k, v in hj["top2bottom"].items()}, hj["none_bottoms"])
    b = synth_batch("bert", 30522, hier, 4096, 5, 128, seed=1)
    L = np.array(b["lens"])
    assert 43 < L.mean() < 50 and L.max() <= 128 and L.min() >= 8          # mean 46.3
    ids, seg = b["ids"].numpy(), b["seg"].numpy()
    assert (ids[:, 0] == 101).all()
    for r in range(64):
        row, sg = ids[r, :L[r]], seg[r, :L[r]]
        fs = int(np.argmax(row == 102))
        assert (sg[:fs] == 0).all() and (sg[fs:] == 1).all() and (ids[r, L[r]:] == 0).all()
    lab = b["labels"].numpy()
    assert 1.2 < lab.sum(1).mean() < 1.45                                   # 1.32 labels / utterance
    none_cols = sorted(hier.none_bottoms)
    assert lab[:, none_cols].sum() == 0                                     # a gold label is never a NONE label
    for k in hier.group_tops:
        assert (lab[:, hier.top2bottom[k]].sum(1) <= 1).all()               # STC_util.py:34 invariant
    x = synth_batch("xlm-roberta", 250002, hier, 8, 5, 128, seed=2)
    assert (x["ids"][:, 0] == 0).all() and int(x["ids"].max()) < 250002
    d = synth_batch("bert", 30522, hier, 8, 5, 128, seed=3, dense=True)
    assert d["lens"] == [128] * 8


def test_hierarchy_tables_are_consistent():
    from oracle import stc_oracle as O
    hj = json.load(open(os.path.join(GOLD, "dstc2_hierarchy.json")))
    hier = O.Hierarchy({int(k): v for k, v in hj["top2bottom"].items()}, hj["none_bottoms"])
    assert (hier.n_top, hier.n_bottom, hier.n_groups, hier.n_cols) == (30, 161, 10, 171)
    assert [len(hier.top2bottom[k]) for k in hier.group_tops] == [75, 27, 4, 5, 7, 6, 4, 7, 4, 2]
    scored = [b for b in hier.col_bottom if b >= 0]
    assert sorted(scored) == list(range(161))                                # every bottom label scored exactly once
    assert sum(hier.none_col) == 10                                          # one NONE value per multi-way group
    for g in range(hier.n_groups):
        assert hier.none_col[hier.grp_off[g + 1] - 1] == 1                   # ... and it is the group's last column


def test_dropout_hash_restatement_is_a_sound_bernoulli_source():
    """The counter-based dropout hash (csrc/ptx.cuh dropout_quad, restated in oracle.stc_oracle.dropout_lanes): uniform
    16-bit lanes, drop rate = p, no correlation between lanes / neighbours / the strides the kernels use, unrelated streams
    for neighbouring seeds, geometric gaps between drops."""
    import numpy as np
    from oracle import stc_oracle as O
    n_q = 1 << 20
    q = np.arange(n_q)
    for p in (0.1, 0.3):
        thr = O.dropout_threshold(p)
        for raw_seed in (1, 2, 5 * 1000003 + 3 * 16 + 2):
            lanes = O.dropout_lanes(O.mix_seed(raw_seed), q)
            drop = lanes < thr
            assert np.all(np.abs(drop.mean(0) - p) < 4 * np.sqrt(p * (1 - p) / n_q))
            for i in range(4):      # chi-square on 256 bins, 255 dof: mean 255, sd 22.6
                h = np.bincount(lanes[:, i] >> 8, minlength=256)
                assert ((h - n_q / 256) ** 2 / (n_q / 256)).sum() < 255 + 6 * 22.6
            d = drop.astype(np.float64) - p
            corr = lambda a, b: abs(float((a * b).mean())) / (p * (1 - p))
            for i in range(4):
                for k in range(i + 1, 4):
                    assert corr(d[:, i], d[:, k]) < 5e-3
            flat = d.reshape(-1)
            for s in (1, 2, 3, 4, 8, 16, 64, 128, 512, 768, 3072, 65536):
                assert corr(flat[:-s], flat[s:]) < 5e-3, s
    a = (O.dropout_lanes(O.mix_seed(7), q) < 6554).astype(float) - 0.1
    b = (O.dropout_lanes(O.mix_seed(8), q) < 6554).astype(float) - 0.1
    assert abs(float((a * b).mean())) / 0.09 < 5e-3
    gaps = np.diff(np.nonzero(O.dropout_keep_mask(3, 1 << 21, 0.1) == 0)[0])
    assert abs(gaps.mean() - 10.0) < 0.1 and abs(gaps.var() - 90.0) < 3.0


def test_trainer_launches_every_bucket_once_at_the_announced_slots():
    """DataParallelTrainer._grad_ready (host logic, no GPU): with collective slots a finished bucket waits for the next
    'slot' the backward announces (or for 'emb') and is launched exactly once, in completion order; the lowest layer's
    bucket — its weight-gradient GEMMs are issued AFTER the embedding backward (model._encoder_backward) — is announced
    behind 'emb' and launched at once; without slots everything is launched immediately; names that are not buckets
    ('head', layers inside a merged bucket) are ignored."""
    from nbest_b200.trainer import DataParallelTrainer
    events = ["head"]
    for l in reversed(range(1, 4)):
        events += ["slot", "layer%d" % l]
    events += ["slot", "emb", "layer0"]

    def run(slots, buckets):
        t = object.__new__(DataParallelTrainer)
        t.comm_slots, t._pending, t._bucket_names, launched = slots, [], set(buckets), []
        t.timeline, t._emb_seen = None, False
        t._launch_bucket = lambda name: launched.append((name, len(seen)))
        seen = []
        for e in events:
            seen.append(e)
            t._grad_ready(e)
        return launched

    one = ["layer3", "layer2", "layer1", "emb", "layer0"]
    got = run(True, one)
    assert [n for n, _ in got] == one
    # layer3's bucket starts at the slot announced before layer2's backward window, ..., emb when it is announced, layer0 last
    pos = {n: i for n, i in got}
    assert events[pos["layer3"] - 1] == "slot" and events[pos["layer1"] - 1] == "slot"
    assert events[pos["emb"] - 1] == "emb" and events[pos["layer0"] - 1] == "layer0" and pos["layer0"] == len(events)
    got = run(False, one)
    assert [n for n, _ in got if n in one] == one and all(events[i - 1] == n for n, i in got)      # immediately
    two = ["layer2", "emb", "layer0"]                                  # two layers per bucket: named after the last to finish
    assert [n for n, _ in run(True, two)] == two


def test_nccl_grad_pool_is_best_effort_without_nccl():
    """trainer.NcclGradPool (gradient buffer in NCCL-registered memory) degrades to ordinary allocation when there is no
    NCCL communicator: nothing raises, `ok` is False with a reason, and FlatBuffers allocates as usual inside `with pool`."""
    from nbest_b200 import optim
    from nbest_b200.trainer import NcclGradPool
    pool = NcclGradPool("cpu")
    assert pool.ok is False and pool.why
    with pool:
        assert optim.GRAD_POOL is None
        fb = optim.FlatBuffers([(4, 8), (8,)], "cpu", with_bf16=False)
    assert optim.GRAD_POOL is None and fb.grads.shape == fb.params.shape and float(fb.grads.abs().sum()) == 0.0
    assert pool.register() is False


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference step on the host cores) prints exactly one JSON line with
    the contract's keys; it is the one bench leg that runs without a GPU."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "utterances/s" and d["value"] > 0 and d["vs_baseline"] is None
    # "reference" = the unmodified reference staged under oracle/_ref (or /root/reference here); "port" only without it
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    from oracle import ref_loader
    if ref_loader.available():
        assert d["cpu_baseline"]["kind"] == "reference" and set(d["cpu_baseline"]["split"]) == {"fwd_ms", "bwd_ms", "optimizer_ms"}
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in d["config"] and "model" not in d["config"]
