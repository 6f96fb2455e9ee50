"""Host-side mirrors of the reference's epoch plumbing (nbest_b200.epoch / .data) against golden vectors produced by the
live reference (tests/golden/epoch_valid48.npz, oracle/make_golden.py epoch_case): collate_fn labels, pred_one_sample
strings, filter_informative, update_f1 / compute_f1 counters; and the pre-tokenised data path against the reference's
prepare_inputs_for_roberta fixture. CPU only."""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mod(name):
    """Load a host-only module of the package without importing the CUDA binding."""
    import nbest_b200  # noqa: F401  (package alias)
    return importlib.import_module("nbest_b200." + name)


def _fixture():
    fx = np.load(os.path.join(GOLD, "epoch_valid48.npz"))
    meta = json.loads(str(fx["meta"]))
    hj = json.load(open(os.path.join(GOLD, "dstc2_hierarchy.json")))
    t2b = {int(k): v for k, v in hj["top2bottom"].items()}
    idx2label = {int(k): v for k, v in hj["idx2label"].items()}
    return fx, meta, hj, t2b, idx2label


def test_collate_labels_matches_reference_collate_fn():
    E = _mod("epoch")
    fx, meta, _, _, _ = _fixture()
    ours = E.collate_labels(meta["label_lists"], meta["label2idx"])
    assert torch.equal(ours, torch.from_numpy(fx["labels"]))
    assert ours[:, E.UNK_LABEL_IDX].sum() == 8          # the 8 utterances with a label outside label2idx


def test_decode_strings_filter_and_f1_match_reference():
    from oracle import stc_oracle as O
    E = _mod("epoch")
    fx, meta, hj, t2b, idx2label = _fixture()
    hier = O.Hierarchy(t2b, hj["none_bottoms"])
    top = torch.from_numpy(fx["top"])
    flat = torch.from_numpy(fx["bottom"])
    bottoms, c = {}, 0
    for k in sorted(t2b):
        if len(t2b[k]) > 1:
            bottoms["lin_%d" % k] = flat[:, c:c + len(t2b[k])]
            c += len(t2b[k])
    bitmap = np.asarray(O.decode(hier, top, bottoms))
    preds = E.decode_to_labels(bitmap, t2b, idx2label)
    assert preds == meta["preds"]                       # same labels in pred_one_sample's order
    mask = E.informative_mask(idx2label, meta["ontology"], hier.n_bottom)
    assert E.decode_to_labels(bitmap * mask[None, :], t2b, idx2label) == meta["preds_filtered"]
    # counters from the bitmaps (what nbest_stc_metrics computes) == the reference's string-set counters
    pred = bitmap > 0
    gold_f = E.collate_labels(meta["label_lists"], meta["label2idx"], ontology=meta["ontology"]).numpy() > 0.5
    for m_, gold, want in ((None, fx["labels"] > 0.5, fx["counts"]), (mask.astype(bool), gold_f, fx["counts_filtered"])):
        p, g = (pred, gold) if m_ is None else (pred & m_[None, :], gold & m_[None, :])
        tp, fp, fn = int((p & g).sum()), int((p & ~g).sum()), int((~p & g).sum())
        exact = int(((p == g).all(1)).sum())
        assert [tp, fp, fn, exact] == want.tolist()
    tp, fp, fn = 0, 0, 0
    for pr, go in zip(meta["preds"], meta["label_lists"]):
        tp, fp, fn = E.update_f1(pr, go, tp, fp, fn)
    assert [tp, fp, fn] == fx["counts"][:3].tolist()
    assert np.allclose(E.compute_f1(tp, fp, fn), fx["prf"]) and E.compute_f1(0, 3, 4) == (0, 0, 0)


def test_pretokenized_dataset_reproduces_reference_padded_tensors(tmp_path):
    """pretokenize -> PretokenizedDataset.batch gives exactly the tensors the reference's prepare_inputs_for_roberta builds
    (packing_valid24 fixture, all four layout variants) and collate_fn's labels."""
    from argparse import Namespace
    sys.path.insert(0, os.path.dirname(__file__))
    from fake_tokenizer import FakeTok, FakeXlmrTok
    D = _mod("data")
    fx = np.load(os.path.join(GOLD, "packing_valid24.npz"))
    raw_in = [s.split(" ") for s in fx["raw_in"].tolist()]
    label2idx = {"<pad>": 0, "<unk>": 1, "a": 2, "b": 3, "c": 4}        # utils/Constants.py: PAD = 0, UNK = 1
    labels = [["a"], [], ["b", "zzz"]] * 8
    cases = (("default", FakeTok(), dict(without_system_act=False, tod_pre_trained_model=None, pre_trained_model="bert")),
             ("nosys", FakeTok(), dict(without_system_act=True, tod_pre_trained_model=None, pre_trained_model="bert")),
             ("tod", FakeTok(), dict(without_system_act=False, tod_pre_trained_model="tod-bert", pre_trained_model=None)),
             ("xlmr", FakeXlmrTok(), dict(without_system_act=False, tod_pre_trained_model=None, pre_trained_model="xlm-roberta")))
    for tag, tok, kw in cases:
        out = D.pretokenize(raw_in, raw_in[::-1], labels, tok, Namespace(**kw), label2idx, str(tmp_path / tag), chunk=7)
        ds = D.PretokenizedDataset(out)
        assert len(ds) == 24
        b = ds.batch(np.arange(24), pinned=False)
        assert np.array_equal(b["ids"].numpy(), fx["ids_" + tag]) and b["lens"] == fx["lens_" + tag].tolist()
        if "seg_" + tag in fx.files:
            assert np.array_equal(b["seg"].numpy(), fx["seg_" + tag])
        else:
            assert b["seg"] is None
        assert np.array_equal(b["trans_ids"].numpy()[::-1][:, :1], fx["ids_" + tag][:, :1])     # reversed stream: same rows
        assert b["labels"].shape == (24, len(label2idx))
        sub = ds.batch([5, 2, 9], pinned=False)                                               # ragged sub-batch: batch-max padding
        S = int(max(fx["lens_" + tag][[5, 2, 9]]))
        assert sub["ids"].shape == (3, S) and np.array_equal(sub["ids"].numpy(), fx["ids_" + tag][[5, 2, 9], :S])
    # labels: unknown strings -> the <unk> column (tod_asr_util.py:119)
    ds = D.PretokenizedDataset(str(tmp_path / "default"))
    lab = ds.batch([2], pinned=False)["labels"][0]
    assert lab.nonzero().flatten().tolist() == [1, 3]


def test_epoch_order_partitions_every_global_batch_across_ranks():
    D = _mod("data")
    n, bs, world = 103, 8, 4
    per_rank = [D.epoch_order(n, bs, True, 5, 2, r, world) for r in range(world)]
    assert len({len(x) for x in per_rank}) == 1
    seen = np.concatenate([np.concatenate(x) for x in per_rank])
    assert len(set(seen.tolist())) == len(seen)                    # disjoint
    for step in range(len(per_rank[0])):
        assert len({len(per_rank[r][step]) for r in range(world)}) == 1
    assert len(seen) >= n - world * bs                             # only a ragged tail is dropped
    a = D.epoch_order(n, bs, True, 5, 2)
    b = D.epoch_order(n, bs, True, 5, 3)
    assert sorted(np.concatenate(a).tolist()) == list(range(n)) and not np.array_equal(a[0], b[0])


def test_prefetcher_yields_the_dataset_batches_in_order_on_cpu(tmp_path):
    from argparse import Namespace
    sys.path.insert(0, os.path.dirname(__file__))
    from fake_tokenizer import FakeTok
    D = _mod("data")
    fx = np.load(os.path.join(GOLD, "packing_valid24.npz"))
    raw_in = [s.split(" ") for s in fx["raw_in"].tolist()]
    opt = Namespace(without_system_act=False, tod_pre_trained_model=None, pre_trained_model="bert")
    ds = D.PretokenizedDataset(D.pretokenize(raw_in, raw_in, [["x"]] * 24, FakeTok(), opt, {"<pad>": 0, "<unk>": 1, "x": 2},
                                             str(tmp_path / "d")))
    batches = D.epoch_order(24, 5, False, 0, 0)
    got = list(D.Prefetcher(ds, batches, "cpu", depth=2))
    assert len(got) == 5
    for b, idx in zip(got, batches):
        ref = ds.batch(idx, pinned=False)
        assert torch.equal(b["ids"], ref["ids"]) and torch.equal(b["seg"], ref["seg"]) and b["lens"] == ref["lens"]
        assert torch.equal(b["labels"], ref["labels"])


def test_coverage_sampler_equals_the_reference_pandas_sampler():
    """`--coverage` (utils/dataset/tod_asr_util.py:12-39): the numpy restatement keeps exactly the utterances the live
    reference's pandas sampler kept on the shipped valid file (fixture: oracle/make_golden.py --coverage-only). The
    sampler only looks at the label lists, so the fixture stores them integer-coded."""
    from nbest_b200.data import stratified_sample
    fx = np.load(os.path.join(GOLD, "coverage_valid.npz"))
    labels = [[str(c)] for c in fx["label_code"]]
    n = len(labels)
    for cov in (0.1, 0.25, 0.5):
        idx = stratified_sample([None] * n, [None] * n, labels, cov)
        assert len(idx) == int(fx["count_%g" % cov][0])
        assert np.array_equal(fx["label_code"][idx], fx["labels_%g" % cov])          # same label sequence, same order
        assert len(set(idx.tolist())) == len(idx)                                      # without replacement
    # first occurrences come first, in file order
    first = stratified_sample([None] * n, [None] * n, labels, 0.1)[:len(set(fx["label_code"].tolist()))]
    assert np.array_equal(first, np.sort(first)) and len(set(fx["label_code"][first].tolist())) == len(first)
    import pytest
    with pytest.raises(ValueError):
        stratified_sample([None] * 4, [None] * 4, [["a"], ["a"], ["b"], ["b"]], 2.0)   # n > population, like pandas
