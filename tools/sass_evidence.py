"""Per-kernel SASS evidence of the built library (no GPU needed): instruction count, registers, and how often the
Blackwell-only mnemonics appear -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor
loads / stores), UBLKCP (bulk copy), FFMA2 / FMUL2 / FADD2 (packed fp32 pairs), MUFU.
usage: python tools/sass_evidence.py > profiles/<tag>_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speechmix_b200", "libspeechmix_sm100.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "MUFU", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    counts, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            counts[name]["n"] += 1
            body = line.split("*/", 1)[1]
            for mn in MNEMONICS:
                if re.search(r"\b" + mn + r"\b", body) or (mn in ("MUFU",) and "MUFU." in body):
                    counts[name][mn] += 1
    names = demangle(list(counts))
    print("# cuobjdump -sass / -res-usage of speechmix_b200/libspeechmix_sm100.so (sm_100a); counts are static SASS occurrences")
    print("%-78s %6s %4s  %s" % ("kernel", "instr", "regs", " ".join("%7s" % m for m in MNEMONICS)))
    for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["n"]):
        short = re.sub(r"\(.*", "", names.get(k, k)).replace("smx::", "").replace("void ", "")
        print("%-78s %6d %4s  %s" % (short[:78], c["n"], regs.get(k, "?"), " ".join("%7d" % c[m] for m in MNEMONICS)))


if __name__ == "__main__":
    sys.exit(main())
