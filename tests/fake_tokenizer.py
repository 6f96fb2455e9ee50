"""Deterministic stand-in for a HuggingFace tokenizer (no vocabularies exist offline). Exposes exactly what
utils/bert_xlnet_inputs.py uses: cls_token, sep_token, pad_token_id, tokenize, convert_tokens_to_ids.
oracle/make_golden.py feeds it to the reference's prepare_inputs_for_roberta; the tests feed it to the drop-in."""


def _fnv(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


class FakeTok:
    cls_token, sep_token, pad_token_id = "[CLS]", "[SEP]", 0

    def tokenize(self, w):                                               # 1-3 deterministic word pieces
        if w in ("[SEP]", "[CLS]"):
            return [w]
        n = 1 + (sum(map(ord, w)) % 3 if len(w) > 4 else 0)
        return [w] if n == 1 else [w[:2]] + ["##" + w[2 + i:3 + i] for i in range(n - 1)]

    def convert_tokens_to_ids(self, toks):
        sp = {"[CLS]": 101, "[SEP]": 102}
        return [sp.get(t, 1000 + (_fnv(t) % 29000)) for t in toks]


class FakeXlmrTok(FakeTok):
    cls_token, sep_token, pad_token_id = "<s>", "</s>", 1

    def tokenize(self, w):
        return [w] if w in ("<s>", "</s>", "</s></s>") else FakeTok.tokenize(self, w)

    def convert_tokens_to_ids(self, toks):
        sp = {"<s>": 0, "</s>": 2, "<pad>": 1, "</s></s>": 3}            # '</s></s>' is one unknown piece (-> <unk> = 3)
        return [sp.get(t, 1000 + (_fnv(t) % 249000)) for t in toks]
