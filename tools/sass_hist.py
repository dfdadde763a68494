"""tools/sass_hist.py -- SASS opcode histogram per kernel of libslq_b200.so (runs in the build container:
cuobjdump needs no GPU).  Shows which kernels carry the Blackwell-native instructions
(UTCIMMA / UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA, LDTM / STTM = tcgen05.ld / st, FFMA2 = packed fp32).

    python tools/sass_hist.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "semilayer-wise-mixed-precision-quantization_b200", "libslq_b200.so")
KEY = ("UTCIMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA2",
       "HMMA", "IMMA", "IDP", "I2F", "F2I", "I2IP", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "BAR")


def main():
    out = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode histogram per kernel of libslq_b200.so (cuobjdump -sass, sm_100a); key opcodes first")
    for (name, cnt), pretty in zip(kernels.items(), demangle):
        total = sum(cnt.values())
        keyed = {}
        for op, n in cnt.items():
            for k in KEY:
                if op.startswith(k):
                    keyed[k] = keyed.get(k, 0) + n
        print("\n%s\n  %d instructions; %s" % (pretty[:200], total, " ".join("%s=%d" % kv for kv in sorted(keyed.items()))))
        print("  top: " + " ".join("%s=%d" % kv for kv in cnt.most_common(12)))


if __name__ == "__main__":
    sys.exit(main())
