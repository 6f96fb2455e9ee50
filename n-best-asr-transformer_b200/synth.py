"""Synthetic DSTC2-shaped batches at the token-id level (no tokenizer vocabularies or datasets exist offline).

Shape statistics are the empirical distributions of the reference's shipped fixture
`dstc2_data/processed_data/raw/valid` (3,560 utterances; measured once in the build container):
words in the last system turn, words per ASR hypothesis, words in the transcript, labels per utterance; word ->
word-piece inflation 1.15. Layout follows utils/bert_xlnet_inputs.py:75-85:
    [CLS] sys ... [SEP] hyp1 [SEP] hyp2 ... hypN [SEP]        segment 0 up to the first [SEP] (exclusive), 1 after
    [CLS] sys ... [SEP] transcript [SEP]                      (transcript stream)
right padded to the batch maximum with pad id 0 (BERT) / 1 (XLM-R); XLM-R uses <s>=0 and </s>=2.
"""
import numpy as np
import torch

_SYS = {4: 0.0062, 5: 0.0326, 6: 0.0427, 7: 0.1025, 8: 0.0427, 9: 0.1025, 10: 0.1354, 11: 0.1098, 12: 0.0416, 13: 0.0295,
        14: 0.0817, 15: 0.0323, 16: 0.059, 17: 0.0281, 18: 0.0166, 19: 0.011, 20: 0.0025, 21: 0.0022, 22: 0.002, 23: 0.0003,
        24: 0.0006, 25: 0.0003, 27: 0.118}
_HYP = {1: 0.0763, 2: 0.1735, 3: 0.2041, 4: 0.1997, 5: 0.1212, 6: 0.0606, 7: 0.039, 8: 0.0269, 9: 0.0229, 10: 0.0181,
        11: 0.0162, 12: 0.0142, 13: 0.0106, 14: 0.0068, 15: 0.0043, 16: 0.0031, 17: 0.0019, 18: 0.0004, 19: 0.0001}
_TR = {1: 0.1742, 2: 0.1795, 3: 0.1093, 4: 0.2497, 5: 0.0831, 6: 0.0489, 7: 0.0323, 8: 0.0298, 9: 0.0247, 10: 0.0121,
       11: 0.0152, 12: 0.0152, 13: 0.011, 14: 0.0065, 15: 0.0034, 16: 0.0039, 17: 0.0008, 18: 0.0003, 22: 0.0003}
_NL = {1: 0.6882, 2: 0.3028, 3: 0.009}
_INFLATE = 1.15


def _draw(rng, table, size):
    ks = np.array(list(table.keys()))
    ps = np.array(list(table.values()), dtype=np.float64)
    return rng.choice(ks, size=size, p=ps / ps.sum())


def _pieces(rng, words):
    """word count -> word-piece count (x1.15 on average, at least one piece per word)."""
    return int(words + rng.binomial(words, _INFLATE - 1.0))


def synth_batch(kind, vocab_size, hier, B, n_hyps=5, max_len=128, seed=999, with_trans=True, dense=False):
    """Returns dict(ids, seg, lens, [trans_ids, trans_seg, trans_lens,] labels) of CPU tensors / lists.

    `hier` needs .n_top, .n_bottom, .top2bottom and .none_bottoms (e.g. ops.DeviceHierarchy).
    dense=True makes every ASR sequence exactly max_len tokens (the worst case quoted for the roofline)."""
    rng = np.random.RandomState(seed)
    if kind == "xlm-roberta":
        cls, sep, pad = 0, 2, 1
    else:
        cls, sep, pad = 101 % vocab_size, 102 % vocab_size, 0
    lo = 1000 if vocab_size > 2000 else 5

    def build(rows_pieces):
        """rows_pieces: per utterance [sys_len, [piece counts of each hypothesis]]"""
        rows, segs = [], []
        for sys_len, hyps in rows_pieces:
            toks, sg = [cls] + list(rng.randint(lo, vocab_size, size=sys_len)), [0] * (1 + sys_len)
            for h in hyps:
                toks += [sep] + list(rng.randint(lo, vocab_size, size=h))
                sg += [1] * (1 + h)
            toks.append(sep)
            sg.append(1)
            rows.append(toks[:max_len])
            segs.append(sg[:max_len])
        lens = [len(r) for r in rows]
        S = max(lens)
        ids = np.full((len(rows), S), pad, dtype=np.int64)
        seg = np.zeros((len(rows), S), dtype=np.int64)
        for i, (r, s) in enumerate(zip(rows, segs)):
            ids[i, :len(r)] = r
            seg[i, :len(s)] = s
        return torch.from_numpy(ids), torch.from_numpy(seg), lens

    sys_words = _draw(rng, _SYS, B)
    asr_rows, tr_rows = [], []
    for b in range(B):
        sys_len = _pieces(rng, int(sys_words[b]))
        # hypotheses of one utterance are variants of the same sentence: a shared log-normal length factor
        # (sigma 0.6, unit mean) reproduces the fixture's spread (5-best: mean 46, p90 76, p99 117 tokens)
        f = np.exp(rng.normal(-0.18, 0.6))
        hyps = [_pieces(rng, int(w)) for w in np.maximum(1, np.round(_draw(rng, _HYP, n_hyps) * f)).astype(int)]
        if dense:
            need = max_len - (2 + sys_len + len(hyps))
            hyps = [max(1, need // len(hyps))] * len(hyps)
            hyps[-1] += max(0, need - sum(hyps))
        asr_rows.append((sys_len, hyps))
        tr_rows.append((sys_len, [_pieces(rng, int(_draw(rng, _TR, 1)[0]))]))
    out = {}
    out["ids"], out["seg"], out["lens"] = build(asr_rows)
    if with_trans:
        out["trans_ids"], out["trans_seg"], out["trans_lens"] = build(tr_rows)
    labels = np.zeros((B, hier.n_bottom), dtype=np.float32)
    none_b = set(getattr(hier, "none_bottoms", ()))
    for b in range(B):
        k = int(_draw(rng, _NL, 1)[0])
        for t in rng.choice(np.arange(2, hier.n_top), size=k, replace=False):   # <= 1 bottom per act-slot group
            cand = [x for x in hier.top2bottom[int(t)] if x not in none_b]
            labels[b, cand[rng.randint(len(cand))]] = 1.0
        if rng.rand() < 0.007:
            labels[b, 1] = 1.0                                                   # '<unk>' label column (0.7 % in the fixture)
    out["labels"] = torch.from_numpy(labels)
    return out
