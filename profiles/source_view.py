#!/usr/bin/env python
"""Per-source-line view of an ncu capture taken with --import-source on.

  python profiles/source_view.py <capture.ncu-rep> <library.so> <kernel substring | "mangled|demangled"> [out.md]

ncu's `--page source --csv` lists warp-state samples per SASS instruction (by address); `nvdisasm --print-line-info` of the same
build lists the CUDA line of every SASS instruction in the same order.  The two streams are aligned by instruction index and
the samples summed per source line (lines of inlined helpers count for the line they live on).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, inside, line = [], False, None
        for ln in txt.splitlines():
            if ln.startswith("\t.section\t.text.") or ln.startswith(".section\t.text."):
                if inside and cur:
                    out.append(cur)
                inside = kernel in ln
                cur, line = [], None
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                cur.append((int(m.group(1), 16), m.group(2).strip(), line))
        if inside and cur:
            out.append(cur)
    return max(out, key=len) if out else []


def ncu_samples(rep, kernel):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    data, hdr, take = [], None, False
    for r in rows:
        if r and r[0] == "Kernel Name":
            take = kernel in r[1]
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if take and hdr and r and r[0].startswith("0x"):
            d = dict(zip(hdr, r))
            data.append((d["Source"].strip(), int(d.get("Warp Stall Sampling (All Samples)", d.get("# Samples", "0")) or 0),
                         int(d.get("Warp Stall Sampling (Not-issued Samples)", "0") or 0), int(d.get("Instructions Executed", "0") or 0)))
    return data


def main():
    rep, so, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    out = sys.argv[4] if len(sys.argv) > 4 else None
    # "<mangled substring>|<demangled substring>" when the two differ (template instantiations)
    k_sass, k_ncu = (kernel.split("|") + [kernel])[:2]
    sass = sass_lines(so, k_sass)
    samp = ncu_samples(rep, k_ncu)
    n = min(len(sass), len(samp))
    per_line, total = {}, 0
    mismatch = 0
    for i in range(n):
        op_a = sass[i][1].split()[0].lstrip("@!UP0123456789 ") if sass[i][1] else ""
        if samp[i][0].split()[:1] != sass[i][1].split()[:1]:
            mismatch += 1
        key = sass[i][2] or ("?", 0)
        e = per_line.setdefault(key, [0, 0, 0])
        e[0] += samp[i][1]
        e[1] += samp[i][2]
        e[2] += samp[i][3]
        total += samp[i][1]
    lines = [f"# source view of `{kernel}` — {rep}", "",
             f"{len(samp)} SASS instructions in the capture, {len(sass)} in the disassembly, {mismatch} opcode mismatches in the aligned prefix; "
             f"{total} warp-state samples.", "", "| file:line | samples | share | not issued | warp instructions |", "|---|---:|---:|---:|---:|"]
    for key, e in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:40]:
        lines.append(f"| {key[0]}:{key[1]} | {e[0]} | {100.0 * e[0] / max(1, total):.1f}% | {e[1]} | {e[2]} |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
