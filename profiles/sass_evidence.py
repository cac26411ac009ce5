#!/usr/bin/env python3
"""Mnemonic counts per kernel of the built library (cuobjdump -sass / -res-usage over every object that goes into
libbvc_b200.so): the evidence that the hot kernels are what DESIGN.md says they are -- UTMALDG (TMA tile loads), SYNCS
(mbarrier), VABSDIFF4 (byte-SIMD SAD), DFMA (fp64 transform), LDGSTS (cp.async), REDUX, ATOMS ... -- and that no
tensor-core instruction is present (SAD is not a contraction; the fp64 transform has a defined fma order).
Usage: python profiles/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "basic_video_codec_b200", "csrc")
COLS = ["UTMALDG", "SYNCS", "VABSDIFF4", "VABSDIFF", "SHF", "DFMA", "DADD", "DMUL", "LDGSTS", "REDUX", "ATOMS", "ATOMG", "SHFL", "BAR",
        "LDS", "STS", "LDG", "STG", "STL", "TENSOR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    res = []
    for n in out:
        n = re.sub(r"bvc::\(anonymous namespace\)::|\(anonymous namespace\)::|^void ", "", n)
        n = re.sub(r"\((CUtensorMap_st|bvc::|unsigned|int|const|long|double|short|float).*$", "", n)
        res.append(n[:72])
    return res


def main():
    subprocess.run(["make", "-s", "-j8", "-C", CSRC], check=True, stdout=subprocess.DEVNULL)
    ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2]
    print(f"# profiles/sass_evidence.py over basic_video_codec_b200/csrc/build/*.o ({ver.strip()}); sm_100a")
    print("# object | kernel | registers | stack | SASS instructions | " + " ".join(COLS) + "   (ATOMG = ATOM/ATOMG/RED, TENSOR = HMMA/IMMA/DMMA/UTCxMMA)")
    for obj in sorted(glob.glob(os.path.join(CSRC, "build", "*.o"))):
        res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
        usage = {}
        cur = None
        for line in res.splitlines():
            m = re.search(r"Function (\S+):", line)
            if m:
                cur = m.group(1)
            m = re.search(r"REG:(\d+) STACK:(\d+)", line)
            if m and cur:
                usage[cur] = (int(m.group(1)), int(m.group(2)))
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernels, name = collections.OrderedDict(), None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = m.group(1)
                kernels[name] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
            if m and name:
                op = m.group(1)
                base = op.split(".")[0]
                c = kernels[name]
                c["_n"] += 1
                c[base] += 1
                if base in ("ATOM", "RED"):
                    c["ATOMG"] += 1
                if base in ("HMMA", "IMMA", "DMMA", "QMMA", "OMMA") or base.startswith("UTC"):
                    c["TENSOR"] += 1
        names = list(kernels)
        for mangled, short in zip(names, demangle(names)):
            c = kernels[mangled]
            reg, stack = usage.get(mangled, (0, 0))
            print(f"{os.path.basename(obj)} | {short} | {reg} | {stack} | {c['_n']} | " + " ".join(str(c[k]) for k in COLS))


if __name__ == "__main__":
    main()
