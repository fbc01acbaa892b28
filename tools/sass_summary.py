"""Per-kernel SASS evidence of the Blackwell-native instructions in libaitb200.so:

    python tools/sass_summary.py > profiles/sass_summary.txt

counts, per kernel, of UTCHMMA (tcgen05.mma; `.2CTA` = cta_group::2), LDTM / STTM (tcgen05.ld / .st, TMEM), UTMALDG / UTMASTG
(TMA tensor loads / stores), UTCBAR (tcgen05.commit), HMMA / IMMA (mma.sync, the pre-Blackwell path), LDGSTS (cp.async),
REDG / RED (global reductions) -- from `cuobjdump -sass`, no GPU needed.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ait_b200", "libaitb200.so")
PATTERNS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "IMMA", "LDGSTS", "REDG", "FFMA2"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for pat in PATTERNS:
            if op.startswith(pat):
                counts[cur][pat] += 1
                break
    names = list(counts)
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n") if names else []
    for n, d in zip(names, dem):
        demangle[n] = d
    total = collections.Counter()
    print("SASS instruction counts per kernel of ait_b200/libaitb200.so (cuobjdump -sass; sm_100a)")
    print("%-110s %s" % ("kernel", "  ".join("%s" % p for p in PATTERNS)))
    for n in names:
        c = counts[n]
        if not any(c.values()):
            continue
        total.update(c)
        name = demangle.get(n, n).replace("(bool)", "").replace("(int)", "")
        name = re.sub(r"\((?:const |CUtensorMap|aitb::|float|int|void|unsigned|long|__nv).*", "", name)[:108]
        print("%-110s %s" % (name, "  ".join("%*d" % (len(p), c[p]) for p in PATTERNS)))
    print("%-110s %s" % ("TOTAL", "  ".join("%*d" % (len(p), total[p]) for p in PATTERNS)))


if __name__ == "__main__":
    sys.exit(main())
