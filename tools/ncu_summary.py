"""Summarise an .ncu-rep: headline metrics + instruction counts by source line.  usage: ncu_summary.py rep [topN]"""
import csv, collections, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); H = rows[0]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'sm__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__sass_thread_inst_executed_op_fadd_pred_on.sum', 'sm__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
for r in rows[2:]:
    print("== kernel", r[H.index('Kernel Name')][:70] if 'Kernel Name' in H else '')
    for w in want:
        if w in H: print(f"  {w:95s} {r[H.index(w)]} {rows[1][H.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
out = collections.defaultdict(lambda: [0, 0]); cur = None; HH = None
for r in csv.reader(src.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": HH = r; continue
    if HH is None: continue
    try: line = int(r[0]); n = int(r[HH.index("Instructions Executed")] or 0); s = int(r[HH.index("# Samples")] or 0)
    except Exception: continue
    out[(cur, line, r[1][:100])][0] += n; out[(cur, line, r[1][:100])][1] += s
items = sorted(out.items(), key=lambda kv: -kv[1][0]); tot = sum(v[0] for _, v in items); stot = sum(v[1] for _, v in items)
print("total warp instructions", tot, "samples", stot)
for (f, l, s), v in items[:topn]:
    print(f"{f[:18]:18s} {l:4d} inst={v[0]/1e6:7.2f}M {v[0]/tot:5.1%} samp={v[1]/max(stot,1):5.1%}  {s.strip()[:84]}")
