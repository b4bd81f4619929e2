#!/usr/bin/env python
"""Attribute the stall samples of an ncu capture of k_step_dmma to CUDA SOURCE LINES.
The SASS page of the report has no line column, so the kernel is disassembled from the very same
libaceqd.so with line info (cuobjdump -xelf + nvdisasm -g) and matched instruction by instruction.
usage: scripts/ncu_lines.py <report.ncu-rep> <mangled-name-substring, e.g. 'Li2ELi1E'> [top N]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "pyaceqd_b200", "csrc", "libaceqd.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
lines_of = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if "k_step_dmma" not in dis:
        continue
    # split into functions
    cur, name, funcs = [], None, {}
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if name:
                funcs[name] = cur
            name, cur = m.group(1), []
            continue
        cur.append(ln)
    if name:
        funcs[name] = cur
    for fn, body in funcs.items():
        if "k_step_dmma" in fn and key in fn:
            seq, line = [], None
            for ln in body:
                m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
                if m:
                    line = (m.group(1), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m:
                    seq.append((line, m.group(2).strip()))
            lines_of = seq
            break
    if lines_of:
        break
if not lines_of:
    sys.exit("kernel not found in the library")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isamp = hdr.index("# Samples")
if len(data) != len(lines_of):
    print(f"warning: {len(data)} profiled instructions vs {len(lines_of)} disassembled (library rebuilt since the capture?)")
agg = collections.Counter()
tot = 0
for (line, _), r in zip(lines_of, data):
    n = int(r[isamp])
    agg[line] += n
    tot += n
print(f"total samples {tot}")
srcs = {}
for (line, n) in agg.most_common(top):
    if line is None:
        print(f"{100.0*n/tot:5.1f}%  <no line>")
        continue
    fn, ln = line
    if fn not in srcs:
        p = os.path.join(root, "pyaceqd_b200", "csrc", fn)
        srcs[fn] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[fn][ln - 1].strip() if 0 < ln <= len(srcs[fn]) else ""
    print(f"{100.0*n/tot:5.1f}%  {fn}:{ln:<5d} {text[:110]}")
