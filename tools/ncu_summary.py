"""Summarise an .ncu-rep (read here, no GPU): per-kernel key metrics + top stall sites.
usage: python tools/ncu_summary.py file.ncu-rep [n_top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
for k, row in enumerate(r[2:]):
    d = dict(zip(hdr, row))
    print(f"==== launch {k}: {d.get('Kernel Name','')[:70]} grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for key in KEYS:
        if key in d:
            print(f"   {key:80s} {d[key]} {units[hdr.index(key)]}")
    st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(d[h]))
          for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and d[h]]
    print("   stalls/issue:", ", ".join(f"{n} {v:.2f}" for n, v in sorted(st, key=lambda x: -x[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for bi, blk in enumerate(blocks[1:]):
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    h = rows[1]
    isrc, iall, iex = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
    scols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    data = []
    for row in rows[2:]:
        try:
            data.append((int(row[iall]), row[isrc].strip(), int(row[iex]), row))
        except Exception:
            pass
    tot = sum(x[0] for x in data) or 1
    ops = collections.Counter()
    for n, s_, ex, row in data:
        t = s_.split()
        if t:
            ops[(t[1] if t[0].startswith('@') else t[0]).split('.')[0]] += ex
    print(f"---- source page, kernel {bi}: {tot} samples; top opcodes by executions:",
          ", ".join(f"{o} {c}" for o, c in ops.most_common(12)))
    for idx, (n, s_, ex, row) in sorted(sorted(enumerate(data), key=lambda x: -x[1][0])[:ntop]):
        stl = {h[i].replace('stall_', ''): row[i] for i in scols if row[i] not in ('0', '')}
        print(f"   {idx:5d} {100*n/tot:5.1f}% ex={ex:9d} {s_[:58]:58s} {stl}")
