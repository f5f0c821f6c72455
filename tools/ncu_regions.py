"""Warp-state samples of one kernel by code region: splits the SASS of an ncu source page
(`ncu -i X.ncu-rep --page source --csv > X.csv`) at barrier instructions and prints, per region, the share of
samples, the warp instructions and the top stall reasons.   usage: python tools/ncu_regions.py X.csv [launch_patterns]"""
import csv
import sys


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(path, npat=1.0):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    i0 = hdr.index("stall_barrier")
    names = hdr[i0:i0 + 17]
    tot = sum(f(r[2]) for r in data)
    print(rows[0][1])
    print("total samples %.0f, SASS instructions %d" % (tot, len(data)))
    cuts = [i for i, r in enumerate(data) if "BAR." in r[1] or "SYNCS.PHASECHK" in r[1]] + [len(data) - 1]
    prev = 0
    for b in cuts:
        seg = data[prev:b + 1]
        s = sum(f(r[2]) for r in seg)
        if s >= 0.005 * tot:
            inst = sum(f(r[5]) for r in seg)
            st = sorted(((sum(f(r[i0 + k]) for r in seg), names[k]) for k in range(17)), reverse=True)[:5]
            print("  SASS %5d-%5d  %5.1f %% of samples  %8.0f warp-instr/pattern  %s" % (
                prev, b, 100 * s / tot, inst / npat, ", ".join("%s %.0f" % (n.replace("stall_", ""), v) for v, n in st)))
        prev = b + 1


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
