"""profiles/ncu_traffic.json <- DRAM bytes per launch of our kernels from ncu --set full captures:
    python scripts/ncu_traffic.py C_ABI_NAME=KERNEL_REGEX:REPORT.ncu-rep ...
A name ending in "+" sums the first launch of every distinct kernel the regex matches (an entry point that is several
kernels, e.g. qec_gen2_bwd+=gen2_bwd_kernel|prep_small|fold_kernel:...)."""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for spec in sys.argv[1:]:
    name, rest = spec.split("=", 1)
    rx, rep = rest.split(":", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    multi = name.endswith("+")
    name = name.rstrip("+")
    seen = {}
    for r in rows[2:]:
        if not re.search(rx, r[hdr.index("Kernel Name")]):
            continue
        if multi:
            kn = r[hdr.index("Kernel Name")]
            if kn in seen:
                continue
            j = hdr.index("gpu__time_duration.sum")
            seen[kn] = (float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")) * UNIT[units[hdr.index("dram__bytes_read.sum")]],
                        float(r[hdr.index("dram__bytes_write.sum")].replace(",", "")) * UNIT[units[hdr.index("dram__bytes_write.sum")]],
                        r[j] + " " + units[j])
            continue
        def val(m):
            i = hdr.index(m)
            return float(r[i].replace(",", "")) * UNIT[units[i]]
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        i = hdr.index("gpu__time_duration.sum")
        out[name] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": r[hdr.index("Kernel Name")][:80],
                     "ncu_duration": r[i] + " " + units[i], "report": os.path.basename(rep)}
        break
    if multi and seen:
        rd, wr = sum(v[0] for v in seen.values()), sum(v[1] for v in seen.values())
        out[name] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": " + ".join(k.split("(")[0][-40:] for k in seen),
                     "ncu_duration": " + ".join(v[2] for v in seen.values()), "report": os.path.basename(rep)}
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
