"""Key metrics of every kernel in an .ncu-rep as JSON (one object per captured launch).
usage: ncu_to_json.py rep out.json [ncu_target_stats.json launch_index]
The optional pair (written by tools/ncu_target.py) adds the focal / candidate / interacting counts of the captured neighbour
launch, from which bench.py derives warp instructions per focal fish."""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); H = rows[0]; U = rows[1]
want = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct", "launch__registers_per_thread": "registers",
        "launch__grid_size": "grid", "sm__warps_active.avg.per_cycle_active": "warps_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum": "ffma", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum": "fadd",
        "sm__sass_thread_inst_executed_op_fmul_pred_on.sum": "fmul", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction"}
def num(v, unit):
    v = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit)
    return v * scale if scale else v
res = []
for r in rows[2:]:
    d = {"kernel": r[H.index("Kernel Name")].split("(")[0].replace("fishk::", "").replace("void ", "")}
    for m, k in want.items():
        if m in H:
            d[k] = num(r[H.index(m)], U[H.index(m)])
    if "dram_read" in d: d["dram_bytes"] = d["dram_read"] + d["dram_write"]
    res.append(d)
if len(sys.argv) > 4:
    st = json.load(open(sys.argv[3]))["per_neighbour_launch"][int(sys.argv[4])]
    for d in res:
        if d["kernel"].startswith("neighbour_kernel"):
            d.update(focal_per_launch=st["focal"], candidates_per_launch=st["candidates"], interacting_per_launch=st["interacting"])
json.dump({"source": rep.split("/")[-1], "units": "duration us, bytes, percent", "launches": res}, open(out, "w"), indent=1)
print(json.dumps(res)[:600])
