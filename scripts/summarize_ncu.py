"""Turns gpurun_out/*.ncu-rep and launch-list CSVs into the tracked text summaries under profiles/.
Usage: python scripts/summarize_ncu.py <tag>      (e.g. r01a)"""
import collections, csv, os, subprocess, sys, re
tag = sys.argv[1]
os.makedirs("profiles", exist_ok=True)
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

def opcode_hist(rep, kernel_regex=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", "regex:" + kernel_regex] if kernel_regex else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    try:
        hdr = rows[1]; ie = hdr.index("Instructions Executed"); src = hdr.index("Source"); sm = hdr.index("# Samples")
    except Exception:
        return None
    op = collections.Counter(); smp = collections.Counter(); tot = 0.0; ts = 0.0
    for r in rows[2:]:
        if len(r) <= max(ie, sm) or r[0] in ("Address", "Kernel Name"): continue
        try: v = float(r[ie] or 0); s = float(r[sm] or 0)
        except ValueError: continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[src]); o = m.group(2).split(".")[0] if m else "?"
        op[o] += v; smp[o] += s; tot += v; ts += s
    return op, smp, tot, ts

def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= v: continue
        name = r[k].split("(")[0]; a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[v])
    return agg

with open(f"profiles/{tag}_summary.txt", "w") as f:
    for name in ("launches_bench.csv", "launches_mel.csv", "launches_cluster.csv"):
        p = os.path.join("gpurun_out", name)
        if not os.path.exists(p): continue
        agg = launches(p); tot = sum(a[1] for a in agg.values())
        f.write(f"== launch list {name} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES) ==\n")
        for n, (c, t) in agg.items():
            f.write(f"  {n:58s} launches={c:4d} total={t/1e6:10.3f} ms  share={t/tot*100:5.1f}%\n")
        f.write("\n")
    for rep in ("prof_mel_f32.ncu-rep", "prof_mel.ncu-rep", "prof_ahc.ncu-rep", "prof_ahc_filter.ncu-rep"):
        p = os.path.join("gpurun_out", rep)
        if not os.path.exists(p): continue
        hdr, units, rows = raw(p)
        f.write(f"== ncu --set full --clock-control none --import-source on : {rep} ==\n")
        for r in rows:
            kn = r[hdr.index("Kernel Name")]
            f.write(f"-- {kn}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m); f.write(f"   {m:90s} {r[i]:>18s} {units[i]}\n")
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
                f.write(f"   traffic = dram read + write = {rd + wr:.3f} {units[hdr.index('dram__bytes_read.sum')]}\n")
            except Exception: pass
        for kr in ({"prof_mel_f32.ncu-rep": ["mel512"], "prof_mel.ncu-rep": ["mel512"], "prof_ahc.ncu-rep": ["ahc_init_nn", "ahc_merge"], "prof_ahc_filter.ncu-rep": ["tile128"]}[rep]):
            h = opcode_hist(p, kr)
            if not h: continue
            op, smp, tot, ts = h
            f.write(f"-- SASS opcode mix of {kr} (warp-level instructions executed; stall samples)\n")
            for o, v in op.most_common(16):
                f.write(f"   {o:10s} {v/tot*100:5.1f}% of {tot:.3e} instr   {smp[o]/max(ts,1)*100:5.1f}% of samples\n")
        f.write("\n")
print(open(f"profiles/{tag}_summary.txt").read())
