"""Recipe: stage the live modules of the UNMODIFIED reference under git-ignored oracle/_ref/ so that they travel to the
GPU box with the repository snapshot (/root/reference does not exist there).

    python oracle/vendor_ref.py            # copy (build container only: needs /root/reference)
    python oracle/vendor_ref.py --check    # report what is staged

TEST INFRASTRUCTURE ONLY. Nothing in the product package imports from oracle/ (tests/test_abi_and_host.py guards that);
oracle/_ref/ is listed in .gitignore (reference sources never enter the history) and NOT in .gpurunignore. Consumers:
`oracle/ref_loader.py` (tests, bench.py's reference / baseline legs). The files are byte-for-byte copies — the two
import stubs the reference needs on this image (no `gpustat`, no `transformers.AdamW`) are installed at import time by
ref_loader, never by editing a copy.

Staged (SURVEY §8(a) cites every one of them): n_best_asr_bert.py, models/{model,optimization}.py,
models/modules/hierarchical_classifier.py, utils/{Constants,STC_util,bert_xlnet_inputs,fscore,gpu_selection,util}.py,
utils/dataset/tod_asr_util.py and the two data fixtures dstc2_data/processed_data/raw/{memory.pt,valid}.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")

FILES = [
    "n_best_asr_bert.py",
    "models/model.py",
    "models/optimization.py",
    "models/modules/hierarchical_classifier.py",
    "utils/Constants.py",
    "utils/STC_util.py",
    "utils/bert_xlnet_inputs.py",
    "utils/fscore.py",
    "utils/gpu_selection.py",
    "utils/util.py",
    "utils/dataset/tod_asr_util.py",
    "dstc2_data/processed_data/raw/memory.pt",
    "dstc2_data/processed_data/raw/valid",
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()[:16]


def stage(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("vendor_ref: %s is absent (GPU box?) — using what is already staged under %s" % (REF, DST))
        return os.path.isdir(DST)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(src) != _sha(dst):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
    with open(os.path.join(DST, "MANIFEST"), "w") as f:
        for rel in FILES:
            f.write("%s  %s\n" % (_sha(os.path.join(DST, rel)), rel))
    if verbose:
        print("vendor_ref: staged %d reference files under %s" % (len(FILES), DST))
    return True


def check():
    ok = True
    for rel in FILES:
        p = os.path.join(DST, rel)
        print("%-50s %s" % (rel, _sha(p) if os.path.exists(p) else "MISSING"))
        ok &= os.path.exists(p)
    return ok


if __name__ == "__main__":
    if "--check" in sys.argv:
        sys.exit(0 if check() else 1)
    sys.exit(0 if stage() else 1)
