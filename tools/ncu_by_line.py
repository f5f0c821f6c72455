"""Attribute ncu per-SASS counters (source page CSV) to CUDA source lines using nvdisasm -g output
of the same function (instructions appear in the same order).
usage: ncu_by_line.py <ncu_sass.csv> <nvdisasm_function.txt>"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] not in ("Address", "Kernel Name")]
# nvdisasm: track current (file,line) -- take the innermost "inlined at" chain start too
cur = None; seq = []
stack = []
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f = m.group(1).split("/")[-1]
        if "inlined at" in m.group(3):
            m2 = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (f, int(m.group(2)), (m2.group(1).split("/")[-1], int(m2.group(2))) if m2 else None)
        else:
            cur = (f, int(m.group(2)), None)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        seq.append((cur, m.group(2)))
print("sass in csv", len(body), "sass in disasm", len(seq))
n = min(len(body), len(seq))
byline = collections.Counter(); sam = collections.Counter(); byfile = collections.Counter()
tot = 0; tots = 0
for i in range(n):
    cnt = int(body[i][ix["Instructions Executed"]]); s = int(body[i][ix["# Samples"]])
    key = seq[i][0][:2] if seq[i][0] else ("?", 0)
    byline[key] += cnt; sam[key] += s; tot += cnt; tots += s
print("total warp instr", tot, "samples", tots)
print("top lines by samples")
for k, v in sam.most_common(40):
    print(f"  {k[0]}:{k[1]:<5d} instr {100*byline[k]/tot:6.2f}%  samples {100*v/tots:6.2f}%")
