"""Development aid: folds the source page of an .ncu-rep (captured with --import-source on, built with -lineinfo)
per CUDA source line: FP64 and other warp-instructions per 32 evaluations (= per warp pass) and the issue-slot
model 2*fp64 + other (an FP64 instruction occupies two issue cycles on this part).
    python tools/fold_ncu_source.py <rep> <evaluations in the launch> [top]"""
import csv, subprocess, sys, collections, os
rep, evals = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
passes = evals / 32.0
fp64_ops = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
per_line = collections.defaultdict(lambda: [0.0, 0.0, 0.0, ""])   # other, fp64, samples, text
per_file = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
cur_file, cur_line, cur_text, hdr = "?", None, "", None
seen = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; i_inst = r.index("Instructions Executed"); i_smp = r.index("# Samples"); continue
    if hdr is None: continue
    if r[0] != "":       # a CUDA source line
        cur_line, cur_text = r[0], r[1].strip(); continue
    sass = r[3].strip()
    if not sass: continue
    if r[2] in seen: continue          # an instruction is listed under every file of its inline stack: keep the first
    seen.add(r[2])
    op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
    num = lambda x: float(x) if x not in ("", "-", "n/a") else 0.0
    n = num(r[i_inst]) / passes
    smp = num(r[i_smp])
    k = (cur_file, cur_line)
    is64 = op.split(".")[0] in fp64_ops
    per_line[k][1 if is64 else 0] += n
    per_line[k][2] += smp
    per_line[k][3] = cur_text
    per_file[cur_file][1 if is64 else 0] += n
    per_file[cur_file][2] += smp
tot_o = sum(v[0] for v in per_file.values()); tot_f = sum(v[1] for v in per_file.values()); tot_s = sum(v[2] for v in per_file.values())
print(f"# per 32 evaluations: other {tot_o:.0f} fp64 {tot_f:.0f}  slots(2*fp64+other) {2*tot_f+tot_o:.0f}  samples {tot_s:.0f}")
for f, v in sorted(per_file.items(), key=lambda kv: -(2 * kv[1][1] + kv[1][0])):
    print(f"{f:28s} other {v[0]:7.1f} fp64 {v[1]:7.1f} samples {100*v[2]/max(tot_s,1):5.1f}%")
print()
for (f, l), v in sorted(per_line.items(), key=lambda kv: -(2 * kv[1][1] + kv[1][0]))[:top]:
    print(f"{f:18s} {l:>4s} other {v[0]:7.1f} fp64 {v[1]:7.1f} slots {2*v[1]+v[0]:7.1f} smp {100*v[2]/max(tot_s,1):4.1f}%  {v[3][:90]}")
