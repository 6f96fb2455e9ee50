"""Drop-in for the reference's utils/bert_xlnet_inputs.py: `prepare_inputs_for_roberta`.

Same signature and return value — (input_ids [B,S] int64 right-padded to the batch maximum, segment ids [B,S] or None,
list of lengths) — so that n_best_asr_bert.py:249-250,322-323 keep working; the string work (word-by-word tokenisation)
is inherently host-side and out of scope for the GPU path. What is new: `pinned=True` returns pinned host tensors for
an asynchronous H2D copy, and the returned lengths let the model pack the batch on the GPU without a device->host sync
(`model(opt, ids, ..., input_lens=lens)`).

Layout (reference :75-85):  [CLS] sys... [SEP] hyp1 [SEP] ... hypN [SEP], segment 0 for `[CLS] sys...`, 1 from the first
separator on. `--without_system_act`: [CLS] hyps [SEP], no segment ids (:70-72). ToD checkpoints: [CLS] [SYS] sys...
[USR] hyps [SEP] (:30-35,55-65). XLM-R: the separator between hypotheses is the single string '</s></s>' (:37-40) and the
first one is appended un-tokenised (:79).
"""
import torch


def prepare_inputs_for_roberta(raw_in, tokenizer, opt, device, pinned=False):
    is_xlmr = bool(getattr(opt, "pre_trained_model", None)) and opt.pre_trained_model == "xlm-roberta"
    tod = bool(getattr(opt, "tod_pre_trained_model", None))
    no_sys = bool(getattr(opt, "without_system_act", False))
    cls, sep = tokenizer.cls_token, tokenizer.sep_token
    hyp_sep = sep + sep if is_xlmr else sep

    def pieces(words):
        out = []
        for w in words:
            out += tokenizer.tokenize(w)
        return out

    rows, segs = [], []
    for words in raw_in:
        u = words.index("[USR]")
        sys_words, usr_words = list(words[2:u]), list(words[u + 1:])      # drops the leading '[CLS] [SYS]' and '[USR]'
        if tod:
            sys_words, usr_words = ["[SYS]"] + sys_words, ["[USR]"] + usr_words
        usr_words = [hyp_sep if w == "[SEP]" else w for w in usr_words]
        a, b = pieces(sys_words), pieces(usr_words)
        if tod:
            a, b = [cls] + a, b + [sep]
        elif no_sys:
            rows.append([cls] + b + [sep])
            continue
        else:
            a, b = [cls] + a, [hyp_sep] + b + [sep]
        rows.append(a + b)
        segs.append([0] * len(a) + [1] * len(b))

    lens = [len(r) for r in rows]
    S = max(lens)
    pad = tokenizer.pad_token_id
    ids = torch.tensor([tokenizer.convert_tokens_to_ids(r) + [pad] * (S - len(r)) for r in rows], dtype=torch.long)
    seg = torch.tensor([s + [0] * (S - len(s)) for s in segs], dtype=torch.long) if segs else None
    if pinned:
        ids = ids.pin_memory()
        seg = seg.pin_memory() if seg is not None else None
        return ids, seg, lens
    dev = torch.device(device)
    ids = ids.to(dev, non_blocking=True)
    seg = seg.to(dev, non_blocking=True) if seg is not None else None
    return ids, seg, lens
