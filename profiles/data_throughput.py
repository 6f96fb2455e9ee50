"""Input pipeline throughput: PretokenizedDataset.batch assembly rate and Prefetcher delivery rate in
utterances/s, against the 26 k utterances/s one B200 consumes (8 x that on a box: one pipeline per rank / process).
Synthetic pre-tokenised directory with the DSTC2 shape statistics of nbest_b200.synth (no tokenizer needed).

    python profiles/data_throughput.py [n_utterances] [batch]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbest_b200.data import PretokenizedDataset, Prefetcher, epoch_order   # noqa: E402
from nbest_b200.synth import synth_batch                                    # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    hj = json.load(open(os.path.join(ROOT, "tests", "golden", "dstc2_hierarchy.json")))
    hier = type("H", (), dict(n_top=len(hj["top2bottom"]), n_bottom=sum(len(v) for v in hj["top2bottom"].values()),
                              top2bottom={int(k): v for k, v in hj["top2bottom"].items()}, none_bottoms=hj["none_bottoms"]))()
    arrays = {k: [] for k in ("asr_ids", "asr_seg", "trans_ids", "trans_seg", "label_idx")}
    lens = {k: [] for k in ("asr", "trans", "label")}
    for s in range(0, n, 2048):
        b = synth_batch("bert", 30522, hier, B=min(2048, n - s), n_hyps=5, max_len=128, seed=s)
        for name, ids, seg, ln in (("asr", b["ids"], b["seg"], b["lens"]), ("trans", b["trans_ids"], b["trans_seg"], b["trans_lens"])):
            ids, seg = ids.numpy(), seg.numpy()
            for i, L in enumerate(ln):
                arrays[name + "_ids"].append(ids[i, :L].astype(np.int32))
                arrays[name + "_seg"].append(int(np.argmax(seg[i, :L] > 0)) if (seg[i, :L] > 0).any() else L)
                lens[name].append(L)
        lab = b["labels"].numpy()
        for i in range(lab.shape[0]):
            idx = np.nonzero(lab[i])[0].astype(np.int32)
            arrays["label_idx"].append(idx)
            lens["label"].append(len(idx))
    d = tempfile.mkdtemp()
    off = lambda ls: np.concatenate([[0], np.cumsum(ls)]).astype(np.int64)
    out = dict(asr_ids=np.concatenate(arrays["asr_ids"]), asr_off=off(lens["asr"]), asr_seg=np.asarray(arrays["asr_seg"], np.int32),
               trans_ids=np.concatenate(arrays["trans_ids"]), trans_off=off(lens["trans"]),
               trans_seg=np.asarray(arrays["trans_seg"], np.int32), label_idx=np.concatenate(arrays["label_idx"]),
               label_off=off(lens["label"]))
    for k, v in out.items():
        np.save(os.path.join(d, k + ".npy"), v)
    json.dump(dict(format="nbest_b200.pretok.v1", n=n, n_labels=hier.n_bottom, pad_token_id=0, has_segment_ids=True), open(os.path.join(d, "meta.json"), "w"))
    ds = PretokenizedDataset(d)
    order = epoch_order(n, bs, True, 999, 0)
    t0 = time.perf_counter()
    for idx in order:
        ds.batch(idx, pinned=False)
    t1 = time.perf_counter()
    print("batch assembly (one thread, numpy gather/scatter, unpinned): %.0f utterances/s" % (n / (t1 - t0)))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    for depth in (3,):
        t0 = time.perf_counter()
        cnt = 0
        for b in Prefetcher(ds, order, dev, depth=depth):
            cnt += b["ids"].shape[0]
        if dev == "cuda":
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        print("Prefetcher -> %s (background assembly, pinned ring, side-stream H2D, depth %d): %.0f utterances/s" % (dev, depth, cnt / (t1 - t0)))


if __name__ == "__main__":
    main()
