#!/usr/bin/env python3
"""Aggregate an ncu report's stall samples of k_rao_fused by kernel PHASE (source-line ranges of raftk_fused.cuh).

usage: tools/ncu_phases.py REPORT.ncu-rep [path/to/libraftk.so matching the report] [kernel substring]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "raft_b200", "csrc", "libraftk.so")
kern = sys.argv[3] if len(sys.argv) > 3 else "k_rao_fusedILi128"
FUSED = "raftk_fused2.cuh" if "fused2" in kern else "raftk_fused.cuh"
srcdir = os.path.dirname(so)
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
# phase markers: find line numbers of comment anchors in the fused source
fsrc = open(os.path.join(srcdir, FUSED)).read().split("\n")
def find(s):
    for i, l in enumerate(fsrc):
        if s in l:
            return i + 1
    return None
marks = [("stage+classes", 1), ("prologue", find("prologue per frequency")), ("part1 walk", find("= pass part 1")),
         ("part1 warp reduce", find("warp sum of the 30 accumulators")), ("cross-warp/cluster reduce", find("for (int t = tid; t < nchunk * 32; t += T) {")),
         ("coefficients+B_drag", find("= linearised coefficients per node")), ("part2 walk", find("= pass part 2")),
         ("assembly", find("double ar[6][6], ai[6][6];")), ("solve6 call+conv", find("const bool ok = solve6")),
         ("flags/cluster sync", find("passes++;"))]
if "fused2" in kern:
    marks = [("stage (TMA blob)", 1), ("prologue", find("---- prologue (a)")), ("part1 walk", find("= pass part 1")),
             ("part1 warp reduce", find("warp sum of the 30 accumulators")), ("cross-warp/cluster reduce", find("for (int t = tid; t < nchunk * 32; t += T) {")),
             ("coefficients+B_drag", find("= linearised coefficients per node")), ("part2 walk", find("= pass part 2")),
             ("park+assembly", find("bin B's drag excitation waits")), ("solve6 call+conv", find("const bool ok = solve6")),
             ("flags/cluster sync", find("passes++;")), ("epilogue", find("if (P.status && rank == 0 && tid == 0)"))]
marks = [(n, l) for n, l in marks if l]
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = [r for r in csv.reader(out.split("\n")) if r]
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: collections.Counter())
tot = totex = 0
kstart = find("k_rao_fused2(DesignsDev D") if "fused2" in kern else find("k_rao_fused(DesignsDev D")
cur_ph, cur_idx, seen_later = marks[0][0], 0, False
for r in rows[2:]:
    try:
        off = int(r[ia], 16) - base
    except Exception:
        continue
    (ln, op) = off2line.get(off, (None, "?"))
    # inlined helpers (proj, solve6, depth_funcs ...) inherit the phase of the surrounding kernel-body code (address order)
    if ln is not None and ln[0] == FUSED and ln[1] >= kstart:
        # phases are laid out in source order; instructions that carry an EARLIER line (rematerialised pointers, hoisted
        # rank / parameter reads) stay with the phase they sit in
        cand = [k for k, (n, l) in enumerate(marks) if l <= ln[1]][-1]
        if cand > 0 or not seen_later:          # stage-phase lines after the prologue began are rematerialised address arithmetic
            cur_idx = cand
        if cand > 0:
            seen_later = True
        cur_ph = marks[cur_idx][0]
    ph = cur_ph
    if ln is not None and ln[0] == "raftk_common.cuh" and ln[1] > 84:
        ph = "solve6 (LU)"
    s, e = int(r[isamp]), int(r[iex])
    agg[ph]["samples"] += s; agg[ph]["instr"] += e; tot += s; totex += e
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU")):
        agg[ph]["fp64"] += e
    for i in stalls:
        try:
            agg[ph][hdr[i]] += int(r[i])
        except ValueError:
            pass
print("total samples %d, warp instructions %d" % (tot, totex))
print("%-28s %8s %8s %8s   top stalls" % ("phase", "samp%", "instr%", "fp64/ins"))
for ph, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted(((k, v) for k, v in c.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:4]
    print("%-28s %7.2f%% %7.2f%% %8.2f   %s" % (ph, 100 * c["samples"] / tot, 100 * c["instr"] / totex, c["fp64"] / max(c["instr"], 1),
                                              ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(c["samples"], 1)) for k, v in st)))
