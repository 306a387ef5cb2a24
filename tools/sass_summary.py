"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (run here, no GPU needed):

  python tools/sass_summary.py > profiles/sass_summary.txt

UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), UTMALDG / UTMASTG = TMA loads / stores (cp.async.bulk.tensor; ".IM2COL" = the
im2col mode), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync (first conv and the producers of the fused
stem kernels only), SYNCS = mbarrier operations."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tensorflow_yolo_b200", "libyolo_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.IM2COL", r"UTMALDG\.\dD\.IM2COL"),
            ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"), ("HMMA", r"\bHMMA"), ("SYNCS", r"\bSYNCS"), ("LDSM", r"\bLDSM")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip() or n
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    total = collections.Counter()
    print("SASS mnemonic counts per kernel of tensorflow_yolo_b200/libyolo_b200.so (cuobjdump -sass, sm_100a)")
    print("{:<96s} {}".format("kernel", " ".join("{:>14s}".format(n) for n, _ in PATTERNS)))
    for fn in order:
        c = counts[fn]
        if not any(c.values()):
            continue
        total.update(c)
        name = demangle(fn)
        cut = name.find(">(")
        name = name[:cut + 1] if cut >= 0 else re.sub(r"\(.*$", "", name)
        name = name.replace("(int)", "").replace("(bool)", "")
        print("{:<96s} {}".format(name[:96], " ".join("{:>14d}".format(c[n]) for n, _ in PATTERNS)))
    print("{:<96s} {}".format("TOTAL", " ".join("{:>14d}".format(total[n]) for n, _ in PATTERNS)))


if __name__ == "__main__":
    main()
