"""Instruction census of libmmvqa_sm100.so per kernel family (what proves a Blackwell-native kernel: UTCHMMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy mma.sync, LDGSTS = cp.async).
    python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mmvqa_b200", "libmmvqa_sm100.so")
PAT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "HMMA", "LDGSTS", "LDSM", "SYNCS", "UCGABAR",
       "MUFU", "ATOMG", "REDG", "RED.E", "ATOM.E", "CCTL"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fam = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        base = re.sub(r"<.*", "", name.replace("void ", "").replace("mmvqa::", ""))
        base = re.sub(r"\(.*", "", base)
        cur = fam.setdefault(base, {"variants": 0, "instr": 0, "counts": collections.Counter()})
        cur["variants"] += 1
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["instr"] += 1
        for p in PAT:
            if op.startswith(p):
                cur["counts"][p] += 1
print("SASS census of %s (cuobjdump -sass), per kernel family: template variants, instructions, marker opcodes" % os.path.basename(LIB))
tot = collections.Counter()
for k, v in fam.items():
    marks = "  ".join("%s %d" % (p, v["counts"][p]) for p in PAT if v["counts"][p])
    print("%-28s variants %3d  instr %7d  %s" % (k, v["variants"], v["instr"], marks))
    tot.update(v["counts"])
print("TOTAL  " + "  ".join("%s %d" % (p, tot[p]) for p in PAT if tot[p]))
