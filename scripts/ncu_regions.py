#!/usr/bin/env python
"""Source-page view of one kernel launch of an ncu report: stall samples, executed warp instructions and lanes per 48-instruction block of
the SASS, written to profiles/<tag>.txt.   python scripts/ncu_regions.py <report.ncu-rep> <kernel-regex:launch-index> <tag>"""
import collections
import csv
import subprocess
import sys

rep, kid, tag = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
S, NS, IE, TIE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
kn = rows[0][1] if rows[0][0] == "Kernel Name" else ""
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) <= TIE or not r[0].strip():
        continue
    try:
        data.append((r[S], int(r[NS] or 0), int(r[IE] or 0), int(r[TIE] or 0)))
    except ValueError:
        pass
tot, toti = sum(d[1] for d in data), sum(d[2] for d in data)
idx_node = next((i for i, d in enumerate(data) if "LDG.E.ENL2.256" in d[0]), -1)
with open("profiles/%s.txt" % tag, "w") as out:
    out.write("# %s\n# launch %s of %s: stall samples and executed instructions per 48-instruction block of the SASS (ncu --page source);\n"
              "# lanes = thread instructions / warp instructions.\n" % (kn[:120], kid, rep))
    out.write("# total: %d samples, %d warp instructions, %d SASS instructions; first 256-bit node load at SASS index %d\n" % (tot, toti, len(data), idx_node))
    out.write("%-12s %9s %9s %6s  %s\n" % ("SASS index", "samples", "warp-inst", "lanes", "memory / control instructions in the block"))
    for i in range(0, len(data), 48):
        seg = data[i:i + 48]
        s, ie, tie = sum(d[1] for d in seg), sum(d[2] for d in seg), sum(d[3] for d in seg)
        ops = [d[0].split()[1] if d[0].startswith("@") else d[0].split()[0] for d in seg if d[0].split()]
        c = collections.Counter(o.split(".")[0] for o in ops if o.startswith(("LDG", "STG", "LDS", "STS", "LDL", "STL", "MUFU", "VOTE", "MATCH", "ATOM", "RED", "CALL", "FFMA", "FMNMX", "PRMT")))
        out.write("%4d-%-7d %8.1f%% %8.1f%% %6.1f  %s\n" % (i, i + len(seg), 100 * s / max(tot, 1), 100 * ie / max(toti, 1), tie / max(ie, 1), ", ".join("%s %d" % kv for kv in c.most_common())))
print(open("profiles/%s.txt" % tag).read())
