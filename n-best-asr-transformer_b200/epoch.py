"""Host-side mirror of the reference's epoch loops around the hot path, with the per-sample host work moved to the device.

Reference (n_best_asr_bert.py): `train_epoch` :232-294, `eval_epoch` :297-388, `pred_one_sample` :198-215,
`filter_informative` :218-229; `update_f1` / `compute_f1` utils/fscore.py:2-21; label multi-hot of `collate_fn`
utils/dataset/tod_asr_util.py:118-126.

What changes against the reference: the loops there call `.tolist()` on the scores and run `pred_one_sample` per sample
(a device->host sync per utterance and per active value group), then compare label-string sets on the host. Here the
head kernel already emits the prediction bitmap (`decode [B,161]`), `nbest_stc_metrics` folds it against the gold
multi-hot into four device counters (TP, FP, FN, exact matches), loss terms stay in a device accumulator, and the host
reads 5 numbers ONCE per epoch. Strings are only materialised when the caller asks for them (prediction dump files).
"""
from collections import OrderedDict

import numpy as np
import torch

from . import ops

UNK_LABEL_IDX = 1      # utils/Constants.py UNK: labels missing from label2idx land in this column (tod_asr_util.py:119)


# ------------------------------------------------------------------------------------------------ labels (A1)
def _informative(lbl, ontology):
    tup = lbl.split("-")
    if len(tup) == 3:
        slot = tup[1]
        return slot == "this" or (slot in ontology["informable"] and len(ontology["informable"][slot]) > 1)
    return True


def collate_labels(label_lists, label2idx, device=None, pinned=False, ontology=None):
    """Multi-hot gold labels [B, len(label2idx)] fp32 (reference collate_fn, tod_asr_util.py:118-126 and :130).
    With `ontology` (evaluation, n_best_asr_bert.py:338-340) a label OUTSIDE label2idx that filter_informative would drop
    is not folded into the <unk> column — known labels are filtered by the column mask of `informative_mask` instead —
    so that the device counters equal the reference's string-set counters exactly."""
    out = torch.zeros(len(label_lists), len(label2idx), dtype=torch.float32)
    for i, labels in enumerate(label_lists):
        for l in labels:
            if l not in label2idx and ontology is not None and not _informative(l, ontology):
                continue
            out[i, label2idx.get(l, UNK_LABEL_IDX)] = 1.0
    if pinned:
        return out.pin_memory()
    return out if device is None else out.to(device, non_blocking=True)


# ------------------------------------------------------------------------------------------------ decode (A11)
def decode_to_labels(decode_rows, top2bottom, idx2label):
    """Label strings of prediction-bitmap rows, in the order pred_one_sample produces them (ascending act-slot id; one
    label per act-slot: its single bottom label, or the arg-max value of its group unless that is the ...-NONE label —
    the kernel already dropped those). decode_rows: [B, n_bottom] uint8 / bool array (host)."""
    rows = np.asarray(decode_rows)
    out = []
    for r in rows:
        labels = []
        for ti in sorted(top2bottom):
            for b in top2bottom[ti]:
                if r[b]:
                    labels.append(idx2label[b])
                    break
        out.append(labels)
    return out


def informative_mask(idx2label, ontology, n_bottom=None):
    """Column mask equivalent to filter_informative (n_best_asr_bert.py:218-229): an `act-slot-value` label is kept iff
    slot == 'this' or the slot is informable with more than one value; labels with fewer parts are always kept."""
    n = n_bottom or (max(idx2label) + 1 if isinstance(idx2label, dict) else len(idx2label))
    keep = np.ones(n, dtype=np.uint8)
    for i in range(n):
        keep[i] = 1 if _informative(idx2label[i], ontology) else 0
    return keep


# ------------------------------------------------------------------------------------------------ f-score (utils/fscore.py)
def update_f1(pred, gold, TP, FP, FN):
    """utils/fscore.py:2-11 on label-string lists (kept for callers that hold strings)."""
    for term in pred:
        if term in gold:
            TP += 1
        else:
            FP += 1
    for term in gold:
        if term not in pred:
            FN += 1
    return TP, FP, FN


def compute_f1(TP, FP, FN):
    """utils/fscore.py:14-21"""
    if TP == 0:
        return 0, 0, 0
    return 100 * TP / (TP + FP), 100 * TP / (TP + FN), 100 * 2 * TP / (2 * TP + FN + FP)


class EpochMetrics:
    """Device-resident epoch accumulators: counters = [TP, FP, FN, exact, utterances] (int64) and the running sum of
    loss_record (= sum of the step's loss terms / batch size, n_best_asr_bert.py:163-195). `result()` is the only
    device->host read."""

    def __init__(self, device, col_mask=None):
        self.counters = torch.zeros(5, dtype=torch.int64, device=device)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=device)
        self.steps = 0
        self.col_mask = None if col_mask is None else torch.as_tensor(np.asarray(col_mask, dtype=np.uint8), device=device)

    def update(self, decode, labels, losses=None):
        """decode [B,n_bottom] uint8 (head kernel), labels [B,n_bottom] fp32 multi-hot, losses [4] device loss terms."""
        ops.stc_metrics(decode, labels, self.counters, self.col_mask)
        self.counters[4] += decode.shape[0]
        if losses is not None:
            self.loss_sum += losses.sum() / decode.shape[0]
        self.steps += 1

    def all_reduce(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.counters, op=dist.ReduceOp.SUM, group=group)
            t = torch.cat([self.loss_sum, torch.tensor([float(self.steps)], device=self.loss_sum.device)])
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return t
        return torch.cat([self.loss_sum, torch.tensor([float(self.steps)], device=self.loss_sum.device)])

    def result(self, group=None):
        """(mean_loss, (p, r, f), acc) — the tuple train_epoch / eval_epoch return (n_best_asr_bert.py:290-294)."""
        t = self.all_reduce(group).tolist()
        tp, fp, fn, exact, tot = self.counters.tolist()
        mean_loss = t[0] / t[1] if t[1] else 0.0
        acc = exact / tot * 100 if tot else 0
        return mean_loss, compute_f1(tp, fp, fn), acc


# ------------------------------------------------------------------------------------------------ epochs (A10 / A12)
def _inputs(batch, opt, prepare, pinned_labels=False):
    batch_labels, raw_in, raw_trans_in, raw_labels = batch
    ids, seg, lens = prepare(raw_in, opt.tokenizer, opt, opt.device)
    tids, tseg, tlens = prepare(raw_trans_in, opt.tokenizer, opt, opt.device)
    if not getattr(opt, "add_segment_ids", False):
        seg = None                                      # n_best_asr_bert.py:252-253: only the ASR stream's ids are dropped
    return batch_labels, raw_in, raw_labels, ids, seg, lens, tids, tseg, tlens


def train_epoch(model, data, opt, memory, trainer=None):
    """Reference train_epoch (n_best_asr_bert.py:232-294): same arguments and return value `(mean_loss, (p, r, f), acc)`.
    `data` yields the reference's collate_fn tuples. Every `opt.n_accum_steps`-th batch takes an optimizer step (:266;
    the reference sets 4 for 12-layer models, builds its loaders with batchSize / 4 and derives t_total from the full
    batchSize, :522-556), the batches in between only accumulate gradients; a trailing incomplete group is dropped by
    the zero_grad at the next epoch start exactly as in the reference (:236). `opt.optim_choice` selects what happens at
    a step (:268-277): bertadam -> step; adam -> global clip at opt.max_norm + step; adamw -> clip + step +
    opt.scheduler.step(). Metrics and losses accumulate on the device.
    `opt.cuda_graphs = True` (one GPU, bertadam) replays the optimizer steps as CUDA graphs (graph.GraphedTrainer): filler
    sequences round each batch's token counts to `opt.graph_bucket = (n_fill, multiple)` so that batches share graphs."""
    from .inputs import prepare_inputs_for_roberta
    from .trainer import DataParallelTrainer
    model.train()
    if trainer is None:
        trainer = getattr(opt, "_nbest_trainer", None)
        if trainer is None or trainer.model is not model or trainer.optimizer is not opt.optimizer:
            trainer = DataParallelTrainer(model, opt.optimizer, add_l2_loss=bool(getattr(opt, "add_l2_loss", False)))
            opt._nbest_trainer = trainer
    opt.optimizer.zero_grad()
    n_accum = max(1, int(getattr(opt, "n_accum_steps", 1)))
    choice = str(getattr(opt, "optim_choice", "bertadam")).lower()
    clip = float(getattr(opt, "max_norm", 0.0)) if choice != "bertadam" else None
    sched = getattr(opt, "scheduler", None) if choice == "adamw" else None
    graphed = None
    if bool(getattr(opt, "cuda_graphs", False)) and choice == "bertadam" and trainer.world == 1:
        from .graph import GraphedTrainer
        graphed = getattr(opt, "_nbest_graphed", None)
        if graphed is None or graphed.trainer is not trainer:
            graphed = GraphedTrainer(trainer, bucket=tuple(getattr(opt, "graph_bucket", (3, 256))),
                                     width=getattr(opt, "graph_width", None))
            opt._nbest_graphed = graphed
    metrics = EpochMetrics(model.device)
    for step, batch in enumerate(data):
        labels, _, _, ids, seg, lens, tids, tseg, tlens = _inputs(batch, opt, prepare_inputs_for_roberta)
        labels = labels.to(model.device, non_blocking=True)
        if (step + 1) % n_accum != 0:
            losses = trainer.accumulate(ids, labels, tids, seg, tseg, lens, tlens)
        elif graphed is not None:
            losses = graphed.step(ids, labels, tids, seg, tseg, lens, tlens)
        else:
            losses = trainer.step(ids, labels, tids, seg, tseg, lens, tlens, clip_norm=clip, scheduler=sched)
        # (a graphed step's head carries the filler rows behind the batch's own: the metrics see the real ones)
        metrics.update(trainer.last_head.decode[:labels.shape[0]], labels, losses)
    return metrics.result()


class EpochInfo:
    """What the reference's EpochInfoCollector carries out of eval_epoch (utils/dataset/tod_asr_util.py:225-242): the
    per-utterance inputs / predictions / golds / match flags and the epoch's summary numbers."""

    def __init__(self, raw_inputs, whole_pred_classes, true_golds, matches, mean_loss, precision, recall, f1, acc):
        self.raw_inputs, self.whole_pred_classes, self.true_golds, self.matches = raw_inputs, whole_pred_classes, true_golds, matches
        self.mean_loss, self.precision, self.recall, self.f1, self.acc = mean_loss, precision, recall, f1, acc


def eval_epoch(model, data, opt, memory, fp=None, efp=None):
    """Reference eval_epoch (n_best_asr_bert.py:297-388): forward + loss (no MSE term, :331) + decode + F1 / accuracy,
    eval mode. Returns what the reference returns: `(mean_loss, (p, r, f), acc, eic)`, or
    `(mean_loss, (p, r, f), acc, all_cases, eic)` when opt.testing (:385-388); `eic` is an EpochInfo (the reference's
    EpochInfoCollector fields). The per-utterance strings (`all_cases`, eic lists, the `fp` / `efp` dump lines :352-357)
    are only materialised when a file object or opt.testing asks for them — that path copies the bitmap to the host once
    per batch; otherwise eic's lists stay empty and everything stays on the device."""
    from .inputs import prepare_inputs_for_roberta
    model.eval()
    want_cases = fp is not None or efp is not None or bool(getattr(opt, "testing", False))
    ontology = getattr(opt, "ontology", None)
    idx2label = memory["idx2label"]
    mask = informative_mask(idx2label, ontology, model.hier.n_bottom) if ontology is not None else None
    metrics = EpochMetrics(model.device, mask)
    cases = []
    with torch.no_grad():
        for batch in data:
            labels, raw_in, raw_labels, ids, seg, lens, tids, tseg, tlens = _inputs(batch, opt, prepare_inputs_for_roberta)
            labels = labels.to(model.device, non_blocking=True)
            losses, head = model.forward_loss_backward(ids, labels, tids, seg, tseg, add_l2_loss=False, input_lens=lens,
                                                       trans_input_lens=tlens, backward=False)
            # the loss sees collate_fn's labels (:331); the filtered metrics see the ontology-aware multi-hot
            gold = labels if ontology is None else collate_labels(raw_labels, memory["label2idx"], model.device,
                                                                  ontology=ontology)
            metrics.update(head.decode, gold, losses)
            if want_cases:
                dec = head.decode.cpu().numpy()
                if mask is not None:
                    dec = dec * mask[None, :]
                preds = decode_to_labels(dec, memory["top2bottom_dict"], idx2label)
                for raw, pred, gold in zip(raw_in, preds, raw_labels):
                    if ontology is not None:
                        gold = [g for g in gold if _informative(g, ontology)]
                    line = "%s\t<=>\t%s\t<=>\t%s\n" % (" ".join(raw), ";".join(pred), ";".join(gold))
                    if fp is not None:
                        fp.write(line)
                    if efp is not None and set(pred) != set(gold):
                        efp.write(line)
                    cases.append((raw, pred, gold))
    mean_loss, prf, acc = metrics.result()
    eic = EpochInfo([" ".join(c[0]) for c in cases], [c[1] for c in cases], [c[2] for c in cases],
                    [set(c[1]) == set(c[2]) for c in cases], mean_loss, prf[0], prf[1], prf[2], acc)
    if bool(getattr(opt, "testing", False)):
        return mean_loss, prf, acc, cases, eic
    return mean_loss, prf, acc, eic


def scores_to_reference_tuple(head, hier):
    """(top_scores, {'lin_k': ...}, final_scores) views of a head result, shaped like model.forward's first three outputs."""
    bottoms = OrderedDict()
    for g, k in enumerate(hier.group_tops):
        c0, c1 = hier.grp_off_host[g] - hier.n_top, hier.grp_off_host[g + 1] - hier.n_top
        bottoms["lin_%d" % k] = head.bottom[:, c0:c1]
    return head.top, bottoms, head.final
