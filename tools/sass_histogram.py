#!/usr/bin/env python
"""Per-kernel histogram of the SASS opcodes that prove the sm_100a path (tcgen05 / TMEM / bulk copies / tensor-map
stores / global reductions), from `cuobjdump -sass` of the built library:

    python tools/sass_histogram.py > profiles/sass_opcodes.txt

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk, UBLKRED = cp.reduce.async.bulk, UTMASTG / UTMALDG = cp.async.bulk.tensor store / load,
HMMA = mma.sync (warp-level tensor-core path of dcn_conv_small.cu), REDG = red.global, SYNCS = mbarrier, LDGSTS = cp.async, ATOMG = atom.global."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jittor_dcn_b200", "libdcn_b200.so")
KEYS = ["UTCHMMA", "HMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UBLKRED", "UTMASTG", "UTMALDG", "UTMAREDG", "REDG", "RED.",
        "ATOMG", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k] += 1
                    break
    names = list(kernels)
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    except (OSError, subprocess.CalledProcessError):
        pass
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: opcode counts per kernel (static instruction counts)")
    for n, c in kernels.items():
        name = demangle.get(n, n)
        name = re.sub(r"\((int|bool|unsigned int)\)", "", name)
        name = re.sub(r"\(.*", "", name)
        cols = " ".join(f"{k.rstrip('.')}={c[k]}" for k in KEYS if c[k])
        print(f"{name}: total={c['_total']} {cols}")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("# library total: " + " ".join(f"{k.rstrip('.')}={tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    sys.exit(main())
