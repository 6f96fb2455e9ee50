"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA
(mma.sync), LDGSTS (cp.async), REDG / RED (fp32 reductions).
usage: python profiles/sass_extract.py > profiles/r2_sass_extract.txt   (runs cuobjdump -sass on the in-tree library)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "n-best-asr-transformer_b200", "libnbest_sm100.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "MUFU", "RED", "ATOM", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    print("# cuobjdump -sass %s  (sm_100a) — instruction counts per kernel" % os.path.relpath(LIB, ROOT))
    print("%-62s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEYS))
    for name, c in counts.items():
        print("%-62s %7d " % (name[:62], c["_total"]) + " ".join("%7d" % c[k] for k in KEYS))


if __name__ == "__main__":
    main()
