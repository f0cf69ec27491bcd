#!/usr/bin/env python3
"""Per-kernel SASS opcode counts of the built library (evidence that the hot kernels use the Blackwell paths they claim).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

Mnemonics (B200_PROFILING.md): DMMA = FP64 tensor pipe (mma.sync.m8n8k4.f64); UTCHMMA = tcgen05.mma kind::f16;
LDTM / STTM = tcgen05.ld / st (tensor memory); UTCBAR = tcgen05.commit; UBLKCP = cp.async.bulk (TMA unit, linear);
SYNCS = mbarrier operations; BAR = named / CTA barriers.
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200", "_lib", "liblipm_mpc.so")
KEYS = ["DMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "BAR", "LDGSTS", "LDS", "STS", "LDG",
        "STG", "DFMA", "DADD", "DMUL", "FFMA", "F2FP", "HADD2", "SHFL", "ATOMG", "RED", "LDL", "STL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(?:\.[A-Za-z0-9_.]+)?[\s;]", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, REPO)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
    print("# columns: total instructions, then the non-zero counts of " + " ".join(KEYS))
    totals = collections.Counter()
    for (mangled, counts), pretty in zip(kernels.items(), demangle):
        pretty = re.sub(r"lipm::\(anonymous namespace\)::", "", pretty)
        pretty = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", pretty)          # drop the parameter list, keep template arguments
        pretty = re.sub(r"\((?:int|bool)\)", "", pretty)
        pretty = re.sub(r"^void ", "", pretty).replace("lipm::<unnamed>::", "")
        row = " ".join(f"{k}={counts[k]}" for k in KEYS if counts[k])
        print(f"{pretty:<58s} {sum(counts.values()):6d}  {row}")
        totals.update(counts)
    print("#")
    print("# library totals: " + " ".join(f"{k}={totals[k]}" for k in KEYS if totals[k]))


if __name__ == "__main__":
    sys.exit(main())
