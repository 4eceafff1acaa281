"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library (cuobjdump -sass):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store /
reduce), UTCBAR (tcgen05.commit), SYNCS (mbarrier).

    python scripts/sass_digest.py > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "midi_emotion_b200", "lib", "libmidi_emotion_b200.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTCBAR|UTCCP|REDG)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)
            counts[kern] = collections.Counter()
            continue
        if kern is None:
            continue
        for mm in PAT.finditer(line):
            key = mm.group(1)
            if key == "UTCHMMA" and ".2CTA" in line:
                key = "UTCHMMA.2CTA"
            counts[kern][key] += 1
    print(f"# {os.path.relpath(LIB, ROOT)} (sm_100a): Blackwell SASS mnemonics per kernel; kernels without any are omitted")
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "REDG"]
    print(f"{'kernel':70s} " + " ".join(f"{c:>12s}" for c in cols))
    tot = collections.Counter()
    for k, c in counts.items():
        if not any(c[x] for x in cols if x != "REDG"):
            continue
        print(f"{k[:70]:70s} " + " ".join(f"{c[x]:12d}" for x in cols))
        tot.update(c)
    print(f"{'TOTAL':70s} " + " ".join(f"{tot[x]:12d}" for x in cols))


if __name__ == "__main__":
    sys.exit(main())
