"""Instruction counts and stall samples of one kernel by the line of the KERNEL BODY they belong to (inlined callees are charged
to the line that called them), from an ncu report with source counters and the cubin's line table.

    python tools/ncu_by_kernel_line.py report.ncu-rep libckm.so <mangled kernel prefix> <source file of the kernel body> [min %]

ncu's own source page charges an inlined instruction to the callee's line AND to every caller line, so its per-line numbers of a
kernel made of inlined helpers do not add up; this tool charges every SASS instruction exactly once."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern, body = sys.argv[1:5]
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.3
innermost = len(sys.argv) > 6 and sys.argv[6] == "inner"  # charge to the innermost frame inside `body` (lambdas of the kernel) instead
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern))
outer = []  # per instruction: line of `body` (outermost frame, or innermost with "inner")
cur = None
in_block = False  # inside a run of "//## File" annotation lines (innermost frame first, then its callers)
fixed = False
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section") or l.startswith(".section"):
        break
    if "//## File" in l:
        if not in_block:
            in_block, fixed = True, False
        for f, n in re.findall(r'File "([^"]+)", line (\d+)', l):
            if f.endswith(body) and not fixed:
                cur = int(n)
                fixed = innermost
        continue
    in_block = False
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        outer.append(cur)
short = re.sub(r"^_ZN\d+[a-z]+(\d+)", "", kern)  # _ZN3ckm15probe_pc_kernelILi31E... -> probe_pc_kernel
short = re.match(r"[A-Za-z_0-9]+?(?=I[LS]|E|$)", short).group(0) if short else kern
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + short],
                        capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hdr = next(r for r in rows if r and r[0] == "Address")
iI, iN = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body_rows = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")][:len(outer)]  # (the first launch that matches)
assert len(body_rows) == len(outer), (len(body_rows), len(outer))
ins, smp = collections.Counter(), collections.Counter()
stalls = collections.defaultdict(collections.Counter)
for r, ln in zip(body_rows, outer):
    ins[ln] += int(r[iI] or 0)
    smp[ln] += int(r[iN] or 0)
    for i, h in stall_cols:
        try:
            stalls[ln][h[6:]] += int(r[i])
        except ValueError:
            pass
ti, ts = sum(ins.values()), sum(smp.values())
src = open(next(os.path.join(d, f) for d, _, fs in os.walk(os.path.dirname(os.path.abspath(lib))) for f in fs if f == os.path.basename(body))).read().split("\n")
print(f"total warp instructions {ti}, stall samples {ts}")
for ln in sorted(k for k in ins if k):
    if 100.0 * ins[ln] / ti >= thr or 100.0 * smp[ln] / max(ts, 1) >= thr:
        top = ", ".join(f"{k} {v}" for k, v in stalls[ln].most_common(3) if v)
        print(f"{ln:5d} {100.0 * ins[ln] / ti:6.2f}% instr {100.0 * smp[ln] / max(ts, 1):6.2f}% samples  {src[ln - 1].strip()[:100]}   [{top}]")
