"""profiles/roofline_traffic.json from an `ncu --set full` capture: dram__bytes_read.sum + dram__bytes_write.sum per launch of
the captured kernel (mean over the captured launches).  bench.py reads that file for roofline.traffic instead of a constant.

    python scripts/ncu_traffic.py gpurun_out/<capture>.ncu-rep "<kernel name as bench.py reports it>" "<provenance note>"
"""
import csv, io, json, os, subprocess, sys

rep, kernel, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def col(name):
    i = hdr.index(name)
    return [float(r[i].replace(",", "")) * scale[units[i]] for r in data]


rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
tot = [a + b for a, b in zip(rd, wr)]
dur_i = hdr.index("gpu__time_duration.sum")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_traffic.json")
try:
    out = json.load(open(path))
except Exception:
    out = {}
out[kernel] = {"dram_bytes_per_launch": sum(tot) / len(tot), "dram_read_bytes": sum(rd) / len(rd), "dram_write_bytes": sum(wr) / len(wr),
               "launches_captured": len(tot), "duration_under_ncu": [f"{r[dur_i]} {units[dur_i]}" for r in data],
               "source": f"{os.path.basename(rep)}: ncu --set full --clock-control none; {note}"}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out[kernel], indent=1))
