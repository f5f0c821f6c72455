"""Summarise an `ncu --page source --csv` dump: instructions by opcode, stall samples by reason."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
op = collections.Counter(); samp = collections.Counter(); stall = collections.Counter()
tot = 0
for r in body:
    if len(r) < len(hdr) or r[0] == "Address" or r[0] == "Kernel Name": continue
    s = r[ix["Source"]].strip()
    toks = s.split()
    o = toks[1] if toks[0].startswith("@") else toks[0]
    o = o.split(".")[0] + ("." + o.split(".")[1] if o.startswith(("LD", "ST", "RED", "ATOM")) and "." in o else "")
    n = int(r[ix["Instructions Executed"]]); tot += n
    op[o] += n
    samp[o] += int(r[ix["# Samples"]])
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[h] += int(r[ix[h]])
print("warp instructions executed:", tot)
for o, n in op.most_common(28):
    print(f"  {o:14s} {n:12d} {100*n/tot:6.2f}%   samples {samp[o]}")
ts = sum(stall.values())
print("stall samples:", ts)
for h, n in stall.most_common(12):
    print(f"  {h:26s} {n:9d} {100*n/ts:6.2f}%")
