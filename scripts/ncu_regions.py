"""aggregate ncu per-line instruction counts by file and by coarse line ranges"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; seen = set(); byfile = collections.Counter(); byline = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "-") and len(r) == len(hdr):
        try: inst = int(r[hdr.index("Instructions Executed")])
        except ValueError: continue
        key = (cur, int(r[0]))
        if key in seen: continue
        seen.add(key); byfile[cur] += inst; byline[key] = inst
tot = sum(byfile.values())
print("total", tot)
for f, v in byfile.most_common(): print(f"{100*v/tot:5.1f}%  {f}")
if len(sys.argv) > 3:
    f = sys.argv[3]; step = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    b = collections.Counter()
    for (ff, ln), v in byline.items():
        if ff == f: b[ln // step * step] += v
    for ln in sorted(b): print(f"   {f}:{ln:4d}-{ln+step-1:<4d} {100*b[ln]/tot:5.1f}%")
