"""Issued warp instructions of one kernel of an .ncu-rep by SASS region: consecutive instructions with the same execution
count are one region (a loop body, a per-item prologue ...).  usage: ncu_sass_regions.py rep kernel_regex [min_share_pct]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) / 100 if len(sys.argv) > 3 else 0.002
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[1]; ia = H.index("Address"); isrc = H.index("Source"); ii = H.index("Instructions Executed"); isamp = H.index("# Samples")
base = int(rows[2][ia], 16)
data = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[ii]), int(r[isamp])) for r in rows[2:] if len(r) > ii and r[ia].startswith("0x")]
tot = sum(d[2] for d in data); stot = sum(d[3] for d in data)
print(rows[0][1]); print("total warp instructions", tot, " samples", stot)
seg = []
for d in data:
    if seg and seg[-1][2] == d[2]: seg[-1][1] = d[0]; seg[-1][3] += 1; seg[-1][4] += d[3]
    else: seg.append([d[0], d[0], d[2], 1, d[3], d[1]])
print("address range   instr  executions      issued   share  samples  first instruction")
for s in seg:
    if s[2] * s[3] > tot * min_share:
        print(f"{s[0]:#06x}-{s[1]:#06x} {s[3]:5d} {s[2]/1e3:10.1f}k {s[2]*s[3]/1e6:9.2f}M {s[2]*s[3]/tot:7.2%} {s[4]/max(stot,1):7.2%}  {s[5][:70]}")
