"""Import the UNMODIFIED reference (TEST INFRASTRUCTURE — only tests/, bench.py's reference / baseline legs and
__graft_entry__ may use this).

The reference tree is looked up in oracle/_ref/ (staged by oracle/vendor_ref.py, travels to the GPU box) and, in the
build container, falls back to /root/reference. Two stubs make it importable on this image without touching a file:
  * `gpustat` (utils/gpu_selection.py:11) is not installed            -> empty module;
  * `transformers.optimization.AdamW` (n_best_asr_bert.py:17) no longer exists in transformers 5.x -> torch.optim.AdamW.
"""
import os
import sys
import types
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref")
LIVE = "/root/reference"


def reference_root():
    if os.path.exists(os.path.join(STAGED, "n_best_asr_bert.py")):
        return STAGED
    if os.path.exists(os.path.join(LIVE, "n_best_asr_bert.py")):
        return LIVE
    return None


def available():
    return reference_root() is not None


_CACHE = {}


def load():
    """-> Namespace(root, nb (module n_best_asr_bert), make_model, BertAdam, reverse_top2bottom, prepare_inputs_for_roberta,
    tod (utils.dataset.tod_asr_util), fscore (utils.fscore))."""
    if "ref" in _CACHE:
        return _CACHE["ref"]
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is not staged: run `python oracle/vendor_ref.py` in the build container")
    if root not in sys.path:
        sys.path.insert(0, root)
    sys.modules.setdefault("gpustat", types.ModuleType("gpustat"))
    import transformers
    import transformers.optimization as topt
    if not hasattr(topt, "AdamW"):
        topt.AdamW = torch.optim.AdamW
        transformers.AdamW = torch.optim.AdamW
    argv, sys.argv = sys.argv, ["x"]
    try:
        import n_best_asr_bert as nb
    finally:
        sys.argv = argv
    from models.model import make_model
    from models.optimization import BertAdam
    from utils.STC_util import reverse_top2bottom
    from utils.bert_xlnet_inputs import prepare_inputs_for_roberta
    import utils.dataset.tod_asr_util as tod
    import utils.fscore as fscore
    ref = Namespace(root=root, nb=nb, make_model=make_model, BertAdam=BertAdam, reverse_top2bottom=reverse_top2bottom,
                    prepare_inputs_for_roberta=prepare_inputs_for_roberta, tod=tod, fscore=fscore)
    _CACHE["ref"] = ref
    return ref


def memory(device=None):
    """The reference's label inventory fixture (dstc2_data/processed_data/raw/memory.pt) + bottom2top_mat as
    n_best_asr_bert.py:500 adds it."""
    ref = load()
    m = torch.load(os.path.join(ref.root, "dstc2_data", "processed_data", "raw", "memory.pt"))
    m["bottom2top_mat"] = ref.reverse_top2bottom(m["top2bottom_dict"])
    if device is not None:
        m["bottom2top_mat"] = m["bottom2top_mat"].to(device)
    return m


def valid_path():
    return os.path.join(load().root, "dstc2_data", "processed_data", "raw", "valid")


def hf_encoder(kind="bert", eager=True, **kw):
    """Random-init HuggingFace encoder of the reference's MODEL_CLASSES (n_best_asr_bert.py:33-37); there are no
    pretrained checkpoints offline. kind: bert | roberta | xlm-roberta."""
    import transformers
    if kind == "bert":
        c = transformers.BertConfig(**kw)
        cls = transformers.BertModel
    elif kind == "roberta":
        d = dict(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1,
                 bos_token_id=0, eos_token_id=2)
        d.update(kw)
        c = transformers.RobertaConfig(**d)
        cls = transformers.RobertaModel
    else:
        d = dict(vocab_size=250002, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1,
                 bos_token_id=0, eos_token_id=2)
        d.update(kw)
        c = transformers.XLMRobertaConfig(**d)
        cls = transformers.XLMRobertaModel
    if eager:
        c._attn_implementation = "eager"
    return cls(c)


def make_opt(encoder, mem, device, pre_trained_model="bert", dropout=0.3, **kw):
    """The fields of the reference's `opt` that the hot path reads."""
    import torch.nn as nn
    opt = Namespace(pretrained_model=encoder, dropout=dropout, device=torch.device(device), score_util="none",
                    sent_repr="cls", cls_type="stc", top2bottom_dict=mem["top2bottom_dict"],
                    label_vocab_size=len(mem["label2idx"]), pre_trained_model=pre_trained_model,
                    tod_pre_trained_model=None, without_system_act=False, add_segment_ids=pre_trained_model == "bert",
                    add_l2_loss=False, optim_choice="bertadam", max_norm=5.0, n_accum_steps=1, ontology=None,
                    testing=False, class_loss_function=nn.BCELoss(reduction="sum"),
                    ce_loss_function=nn.NLLLoss(reduction="sum"), mse_loss_function=nn.MSELoss())
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt
