"""per-source-line instruction counts / stall samples from an .ncu-rep (cuda,sass view)"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
by = 1 if (len(sys.argv) > 4 and sys.argv[4] == "samples") else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; lines = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "-") and len(r) == len(hdr):
        try:
            inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
        except ValueError: continue
        key = (cur_file, int(r[0]))
        if key in lines: continue      # template instantiations repeat the file
        lines[key] = (inst, samp, r[1].strip()[:90])
tot = sum(v[0] for v in lines.values()); tots = sum(v[1] for v in lines.values())
print("total warp-instructions", tot, "samples", tots)
for (f, ln), (inst, samp, src) in sorted(lines.items(), key=lambda kv: -kv[1][by])[:topn]:
    print(f"{100*inst/tot:5.1f}% inst {100*samp/max(tots,1):5.1f}% smp  {f}:{ln:<4d} {src}")
