"""profiles/traffic.json from an `ncu --set full` capture of the seeding kernels (binned seeding: hash_kernel,
count_kernel, bin_prefix_kernel, scatter_sorted_kernel, filter_kernel, seed_kernel of ONE batch).
usage: make_traffic.py capture.ncu-rep pairs_in_capture mode [out.json]
Writes DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per kernel and per pair; bench.py scales the
per-pair figure to its batch for roofline.traffic (the capture must be of the same batch size: the record
traffic of the binned prefilter is not linear in the batch)."""
import csv, io, json, os, subprocess, sys
rep, pairs, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out_path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(r, name):
    i = hdr.index(name)
    v = float(r[i])
    u = units[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
kern, seen = {}, set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
    if name in seen:
        continue  # first launch of every kernel = the first batch
    seen.add(name)
    i_t = hdr.index("gpu__time_duration.sum")
    ms = float(r[i_t]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[units[i_t]]
    kern[name] = {"dram_bytes_read": col(r, "dram__bytes_read.sum"), "dram_bytes_write": col(r, "dram__bytes_write.sum"),
                  "ms_under_ncu": ms, "l2_hit_pct": float(r[hdr.index("lts__t_sector_hit_rate.pct")])}
SEEDING = ("hash_kernel", "count_kernel", "bin_prefix_kernel", "scatter_sorted_kernel", "scatter_kernel", "filter_kernel", "seed_kernel")
other = {k: v for k, v in kern.items() if k not in SEEDING}
kern = {k: v for k, v in kern.items() if k in SEEDING}
total = sum(k["dram_bytes_read"] + k["dram_bytes_write"] for k in kern.values())
try:
    doc = json.load(open(out_path))
    if "kernel" in doc:  # the round-1 layout (one kernel): keep it under its own key
        doc = {"round1_seed_kernel": doc}
except Exception:
    doc = {}
doc[mode] = {"kernel": "seeding (binned): " + " + ".join(kern), "source": os.path.basename(rep) + " (ncu --set full --clock-control none)",
             "pairs_in_capture": pairs, "units_per_pair": 1.0, "kernels": kern, "dram_bytes": total, "dram_bytes_per_pair": total / pairs,
             "other_kernels_of_the_step": other}
json.dump(doc, open(out_path, "w"), indent=1)
print(json.dumps(doc[mode], indent=1))
