"""Parse the PTZ_SETUP_DEBUG / PTZ_TIMING lines of a bench.py run (stderr) and print the slowest set-ups by phase."""
import re
import sys

rows, cur = [], {}
for line in open(sys.argv[1]):
    m = re.match(r"\[ptzba setup\] (.+?)\s+([\d.]+) ms", line)
    if m:
        cur[m.group(1).strip()] = float(m.group(2))
        continue
    m = re.match(r"\[ptzba_solve rank 0\] create ([\d.]+) ms\s+run ([\d.]+) ms", line)
    if m:
        cur["create"], cur["run"] = float(m.group(1)), float(m.group(2))
        rows.append(cur)
        cur = {}
rows = rows[-int(sys.argv[2]):]
keys = ["stream + clock", "orderings (obs)", "block pattern + pair lists", "CG shape + arena", "parameters + work buffers", "create", "run"]
for r in sorted(rows, key=lambda r: -r["create"])[:8]:
    print([(k[:12], r.get(k)) for k in keys])
print("median create", sorted(r["create"] for r in rows)[len(rows) // 2])
