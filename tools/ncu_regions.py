#!/usr/bin/env python3
"""Aggregate an ncu report's per-instruction stall samples by source line / kernel region.

usage: tools/ncu_regions.py REPORT.ncu-rep KERNEL_MANGLED_SUBSTRING [top_n]
Needs the in-tree libraftk.so built with -lineinfo (maps SASS offsets to raft_b200/csrc source lines via nvdisasm)."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "raft_b200", "csrc", "libraftk.so")
SRCDIR = os.path.join(ROOT, "raft_b200", "csrc")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], text=True).split("\n")
infn, cur, off2line = False, None, {}
for l in sass:
    if l.startswith("//--------------------- .text."):
        infn = kern in l
    if not infn:
        continue
    m = re.search(r'//## File ".*?/csrc/([\w.]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S+)", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(out.split("\n")))
rows = [r for r in rows if r]
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(rows[2][ia], 16)
per, perex, ops = collections.Counter(), collections.Counter(), collections.Counter()
tot = totex = 0
for r in rows[2:]:
    try:
        off = int(r[ia], 16) - base
    except Exception:
        continue
    ln, op = off2line.get(off, (None, "?"))
    s, e = int(r[isamp]), int(r[iex])
    per[ln] += s; perex[ln] += e; tot += s; totex += e
    ops[op.split(".")[0]] += e
_src = {}


def src_line(key):
    if not key:
        return ""
    f, n = key
    if f not in _src:
        try:
            _src[f] = open(os.path.join(ROOT, "raft_b200", "csrc", f)).read().split("\n")
        except OSError:
            _src[f] = []
    return _src[f][n - 1].strip()[:110] if 0 < n <= len(_src[f]) else ""
print("total samples", tot, "total warp instructions", totex)
print("opcode mix (executed):", ", ".join("%s %.1f%%" % (k, 100 * v / totex) for k, v in ops.most_common(18)))
print("top lines (samples%, instr%):")
for ln, s in per.most_common(topn):
    print("%-22s %6.2f%% %6.2f%%  %s" % ("%s:%d" % ln if ln else "-", 100 * s / tot, 100 * perex[ln] / totex, src_line(ln)))
