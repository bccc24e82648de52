#!/usr/bin/env python
"""SASS mnemonics of the tensor-core kernels in a built library (no GPU needed):
  python profiles/sass_summary.py redisearch_b200/lib/libvecsim_b200.so [out.md]
The mnemonics that prove tcgen05 / TMEM / TMA (B200_PROFILING.md): UTCHMMA / UTCIMMA (tcgen05.mma kind::f16|tf32 / kind::i8),
.2CTA (cta_group::2), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM alloc), UTMALDG (TMA tensor load),
UBLKCP (cp.async.bulk), .MULTICAST, SYNCS (mbarrier), UCGABAR (cluster barrier)."""
import collections
import re
import subprocess
import sys

KEEP = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT",
        "LDGSTS", "HMMA", "IMMA")


def main():
    so = sys.argv[1]
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    dem = {}
    lines = [f"# SASS of the tensor-core kernels in `{so}` (`cuobjdump -sass`, sm_100a)", "",
             "| kernel | SASS instructions | tensor / TMEM / TMA / barrier mnemonics (count in the instruction stream) |", "|---|---:|---|"]
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M))
        sel = {k: v for k, v in ops.items() if k.split(".")[0] in KEEP}
        if not any(k.startswith(("UTC", "LDTM", "UTMALDG", "UBLKCP")) for k in sel):
            continue
        if name not in dem:
            dem[name] = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        lines.append(f"| `{dem[name]}` | {sum(ops.values())} | " + ", ".join(f"{k} x{v}" for k, v in sorted(sel.items())) + " |")
    out = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(out)
    print(out)


if __name__ == "__main__":
    main()
