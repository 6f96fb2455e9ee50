"""CPU: the prepare_inputs_for_roberta drop-in against the reference's own outputs (tests/golden/packing_valid24.npz,
produced by utils/bert_xlnet_inputs.py on 24 real lines of the shipped `valid` file with tests/fake_tokenizer.py)."""
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(HERE, "golden")

CASES = {
    "default": ("bert", dict(without_system_act=False, tod_pre_trained_model=None, pre_trained_model="bert")),
    "nosys": ("bert", dict(without_system_act=True, tod_pre_trained_model=None, pre_trained_model="bert")),
    "tod": ("bert", dict(without_system_act=False, tod_pre_trained_model="tod-bert", pre_trained_model=None)),
    "xlmr": ("xlmr", dict(without_system_act=False, tod_pre_trained_model=None, pre_trained_model="xlm-roberta")),
}


@pytest.mark.parametrize("tag", sorted(CASES))
def test_prepare_inputs_matches_reference_bit_exact(tag):
    from fake_tokenizer import FakeTok, FakeXlmrTok
    from nbest_b200.inputs import prepare_inputs_for_roberta
    fx = np.load(os.path.join(GOLD, "packing_valid24.npz"))
    raw_in = [str(s).split(" ") for s in fx["raw_in"]]
    tok_kind, kw = CASES[tag]
    tok = FakeXlmrTok() if tok_kind == "xlmr" else FakeTok()
    ids, seg, lens = prepare_inputs_for_roberta(raw_in, tok, Namespace(**kw), torch.device("cpu"))
    assert ids.dtype == torch.long and np.array_equal(ids.numpy(), fx["ids_" + tag])
    assert lens == fx["lens_" + tag].tolist()
    if "seg_" + tag in fx.files:
        assert np.array_equal(seg.numpy(), fx["seg_" + tag])
    else:
        assert seg is None                                   # --without_system_act: no segment ids (reference :70-72)
    if tag == "xlmr":
        assert (ids[:, 0] == 0).all() and int(ids[0, -1]) == 2 and (ids == 3).any()     # <s>, </s>, '</s></s>' -> <unk>


def test_pinned_variant_returns_host_tensors_and_lengths():
    from fake_tokenizer import FakeTok
    from nbest_b200.inputs import prepare_inputs_for_roberta
    raw = [["[CLS]", "[SYS]", "hello", "there", "[USR]", "cheap", "food", "[SEP]", "cheap", "foot"]]
    opt = Namespace(without_system_act=False, tod_pre_trained_model=None, pre_trained_model="bert")
    if not torch.cuda.is_available():
        ids, seg, lens = prepare_inputs_for_roberta(raw, FakeTok(), opt, "cpu")
    else:
        ids, seg, lens = prepare_inputs_for_roberta(raw, FakeTok(), opt, "cuda", pinned=True)
        assert ids.is_pinned() and seg.is_pinned()
    assert lens == [ids.shape[1]] and int(ids[0, 0]) == 101 and int(ids[0, -1]) == 102
    first_sep = ids[0].tolist().index(102)
    assert seg[0].tolist() == [0] * first_sep + [1] * (ids.shape[1] - first_sep)
