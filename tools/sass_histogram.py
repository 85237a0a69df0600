#!/usr/bin/env python
"""Opcode evidence for the tensor kernels: `cuobjdump -sass` of the in-tree libbokego_b200.so, per kernel the counts of the
Blackwell-specific SASS mnemonics (tcgen05 MMA = UTCHMMA / UTCMMA, TMEM load = LDTM, tensor-map TMA = UTMALDG, bulk copy =
UBLKCP, tcgen05.commit = UTCBAR, mbarrier = SYNCS, cluster barrier = UCGABAR, elect = ELECT) next to the warp-level MMA
(HMMA / IMMA) and FFMA counts.  Needs no GPU.  Writes a markdown table to stdout:

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "bokego_b200", "libbokego_b200.so")
WATCH = ["UTCHMMA", "UTCMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR",
         "ELECT", "HMMA", "IMMA", "FFMA", "LDS", "STS", "LDG", "STG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            kernels[cur][m.group(1)] += 1
            if m.group(1) in ("UTCHMMA", "UTCMMA", "UTMALDG", "UTCBAR", "LDTM", "UBLKCP"):
                kernels[cur][m.group(1) + m.group(2)] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    sha = subprocess.run(["sha256sum", SO], capture_output=True, text=True).stdout.split()[0][:16]
    print(f"# SASS opcode histogram of `bokego_b200/libbokego_b200.so` (sha256 {sha}…), `cuobjdump -sass`, sm_100a\n")
    print("Counts of instructions per kernel; Blackwell-only mnemonics first (UTCHMMA / UTCMMA = `tcgen05.mma` kind::f16 / kind::tf32, "
          "LDTM = `tcgen05.ld`, UTMALDG = tensor-map `cp.async.bulk.tensor`, UBLKCP = `cp.async.bulk`, UTCBAR = `tcgen05.commit`, "
          "SYNCS = mbarrier ops, UCGABAR = cluster barrier).\n")
    cols = [c for c in WATCH if any(k[c] for k in kernels.values())]
    print("| kernel | total | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k, c in kernels.items():
        short = re.sub(r"\(.*", "", names[k].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", ""))
        print(f"| `{short}` | {sum(v for n, v in c.items() if '.' not in n)} | " + " | ".join(str(c[x]) for x in cols) + " |")
    print("\n## Modifiers seen on the tensor / TMA instructions\n")
    for k, c in kernels.items():
        mods = sorted(n for n in c if "." in n)
        if mods:
            short = re.sub(r"\(.*", "", names[k].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", ""))
            print(f"- `{short}`: " + ", ".join(f"`{m}` x{c[m]}" for m in mods))


if __name__ == "__main__":
    sys.exit(main())
