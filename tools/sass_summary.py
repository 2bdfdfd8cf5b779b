# SPDX-License-Identifier: Apache-2.0
"""Per-kernel SASS opcode counts of the shipped library (cuobjdump -sass), the evidence that the
hot kernels use the Blackwell tensor / TMA / TMEM paths:  python tools/sass_summary.py > profiles/r2_sass_opcodes.md
  UTCHMMA / UTCQMMA .. = tcgen05.mma        LDTM / STTM = tcgen05.ld / st (TMEM)
  UTCBAR = tcgen05.commit                    UBLKCP = cp.async.bulk (1-D TMA)
  UTMALDG = cp.async.bulk.tensor (TMA tile)  LDGSTS = cp.async (LSU gather)
  SYNCS = mbarrier ops                       RED / REDG = red.global.add   HMMA = legacy mma.sync (must be 0)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "warpconvnet_b200", "csrc", "libwcn_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS",
       "SYNCS", "REDG", "RED", "ATOMG", "HMMA", "IMMA", "LDG", "STG", "LDS", "STS", "SHFL", "MATCH"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            per[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[name]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    per[name][o] += 1
                    break
    print("# SASS opcode counts per kernel of `libwcn_b200.so` (sm_100a; `cuobjdump -sass`, "
          "`tools/sass_summary.py`)\n")
    cols = [o for o in OPS if any(c[o] for c in per.values())]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for k, c in per.items():
        print(f"| `{k[:90]}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in cols) + " |")
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("\nTotals: " + ", ".join(f"{o} {tot[o]}" for o in cols))
    print(f"\nHMMA (legacy mma.sync) instructions in the library: {tot['HMMA']}")


if __name__ == "__main__":
    main()
