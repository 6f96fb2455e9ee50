"""Full training checkpoints: weights + BertAdam moments + step counters + dropout stream position + data cursor.

The reference can only save weights (`save_model` / `load_model`, models/model.py:75-83: `torch.save(state_dict())`, used
for best-on-valid-F1 selection at n_best_asr_bert.py:427-433) and cannot resume an interrupted run. The
weight part written here IS that format — `checkpoint["model"]` has the reference's keys, fp32, `nn.Linear` layout, so
`torch.load(path)["model"]` loads into the reference model and reference checkpoints load here — and next to it the
state needed for a bit-exact continuation of the data-parallel trainer:

  optimizer.m / .v   the flat fp32 Adam moments (models/optimization.py:262-267 `next_m`, `next_v` of every tensor)
  optimizer.steps    per-tensor step counts (the warm-up schedule position, :269-283)
  model.step_seed    position of the counter-based dropout stream
  cursor             caller-defined data position (epoch, batch index, best F1 ...)

Every rank holds identical replicas, so rank 0 writes and every rank reads.
"""
import os

import torch

FORMAT = "nbest_b200.checkpoint.v1"


def save_checkpoint(path, model, optimizer=None, cursor=None):
    """Atomic write (tmp file + rename). Synchronises the device once."""
    ck = dict(format=FORMAT, model={k: v.detach().cpu() for k, v in model.state_dict().items()},
              step_seed=int(model._step_seed), cursor=dict(cursor or {}))
    if optimizer is not None:
        flat = optimizer.flat
        flat.ensure_moments()
        ck["optimizer"] = dict(m=flat.m.detach().cpu(), v=flat.v.detach().cpu(), steps=list(optimizer._steps),
                               total=int(flat.total), offsets=list(flat.offsets),
                               groups=[{k: (list(v) if isinstance(v, tuple) else v) for k, v in g.items() if k != "params"}
                                       for g in optimizer.param_groups])
    tmp = path + ".tmp"
    torch.save(ck, tmp)
    os.replace(tmp, path)
    return path


def load_checkpoint(path, model, optimizer=None, strict=True):
    """Restores weights (and the bf16 working copy), Adam moments, step counters and the dropout stream position.
    Accepts a bare reference `state_dict` file as well (weights only). Returns the saved cursor dict."""
    ck = torch.load(path, map_location="cpu", weights_only=True)
    if not (isinstance(ck, dict) and ck.get("format") == FORMAT):
        model.load_state_dict(ck, strict=strict)            # reference models/model.py:78-83 format
        return {}
    model.load_state_dict(ck["model"], strict=strict)
    model._step_seed = int(ck.get("step_seed", model._step_seed))
    if optimizer is not None and "optimizer" in ck:
        o = ck["optimizer"]
        flat = optimizer.flat
        if int(o["total"]) != int(flat.total) or list(o["offsets"]) != list(flat.offsets):
            raise ValueError("checkpoint optimizer layout (%d elements) does not match this model's flat buffer (%d)"
                             % (int(o["total"]), int(flat.total)))
        flat.ensure_moments()
        flat.m.copy_(o["m"])
        flat.v.copy_(o["v"])
        if len(o["steps"]) != len(optimizer._steps):
            raise ValueError("checkpoint has %d optimizer tensors, the optimizer %d" % (len(o["steps"]), len(optimizer._steps)))
        optimizer._steps = [int(s) for s in o["steps"]]
    return ck.get("cursor", {})
