"""Turn the ncu reports pulled back in gpurun_out/ into the text summaries committed under profiles/.
usage: python tools/summarise_profiles.py <tag>      (e.g. r01)"""
import csv, io, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active"]


def raw(rep):
    rows = list(csv.reader(open(rep)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    return hdr, units, body


def summarise(rep, dst, note):
    hdr, units, body = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# %s\n# source: ncu --set full --clock-control none (%s, exported on the GPU box with --page raw --csv)\n" % (note, os.path.basename(rep)))
        for r in body:
            f.write("\n%s  grid %s block %s\n" % (r[ix["Kernel Name"]], r[ix["Grid Size"]], r[ix["Block Size"]]))
            for k in KEYS:
                if k in ix:
                    f.write("  %-70s %s %s\n" % (k, r[ix[k]], units[ix[k]]))
    return hdr, units, body


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    tot = collections.Counter(); cnt = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]][:90]
        tot[name] += float(r[ix["Metric Value"]].replace(",", "")); cnt[name] += 1
    s = sum(tot.values())
    with open(dst, "w") as f:
        f.write("# launch list of `python bench.py --steps 3 --no-cg` (ncu --metrics gpu__time_duration.sum, "
                "cold-cache, serialised: compare SHARES)\n")
        for k, v in tot.most_common():
            f.write("%10.3f ms %6.2f%% x%4d  %s\n" % (v / 1e6, 100 * v / s, cnt[k], k))


g = os.path.join(ROOT, "gpurun_out"); p = os.path.join(ROOT, "profiles")
if os.path.exists(os.path.join(g, "launches_bench.csv")):
    launches(os.path.join(g, "launches_bench.csv"), os.path.join(p, tag + "_bench_launch_shares.txt"))
for rep, note in (("all128", "every fused pass at 128^2 (tools/prof.py 128 4: 4 angles x 1024 positions)"),
                  ("all256", "fused passes at 256^2 (tools/prof.py 256 1)"),
                  ("bench_grad", "k_grad inside bench.py (c2, 8 angles x 1024 positions): the roofline.traffic source"),
                  ("bench_grad_c4", "k_grad inside bench.py --workload c4 (2 angles x 1024 positions, 256^2): the roofline.traffic source")):
    path = os.path.join(g, rep + "_raw.csv")
    if os.path.exists(path):
        hdr, units, body = summarise(path, os.path.join(p, "%s_%s_ncu.txt" % (tag, rep)), note)
        if rep in ("bench_grad", "bench_grad_c4") and body:
            ix = {h: i for i, h in enumerate(hdr)}
            def mb(k):
                v = float(body[0][ix[k]]); u = units[ix[k]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
            t = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
            tj = os.path.join(p, "traffic.json")
            cur = json.load(open(tj)) if os.path.exists(tj) else {}
            cur["c2" if rep == "bench_grad" else "c4"] = t
            json.dump(cur, open(tj, "w"))
            print("traffic", rep, "bytes/launch:", t)
print("done")
