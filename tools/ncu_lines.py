"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by source line.
usage: ncu_lines.py prof_cs.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = None
fname = ''
lines = {}
cur = None
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples"); i_ie = hdr.index("Instructions Executed")
        i_long = hdr.index("stall_long_sb"); i_short = hdr.index("stall_short_sb"); i_wait = hdr.index("stall_wait")
        i_ni = hdr.index("stall_no_inst")
        i_loc = hdr.index("L2 Theoretical Sectors Local"); i_glob = hdr.index("L2 Theoretical Sectors Global")
        continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    if r[0] != "":  # source line row (aggregated)
        try:
            ln = int(r[0])
        except ValueError:
            continue
        if ln == 0:
            continue
        if not r[i_s].strip().isdigit():
            continue
        key = (fname, ln)
        d = lines.setdefault(key, dict(src=r[1], s=0, ie=0, long=0, short=0, wait=0, loc=0, glob=0, noinst=0))
        d["noinst"] += int(r[i_ni])
        d["s"] += int(r[i_s]); d["ie"] += int(r[i_ie]); d["long"] += int(r[i_long]); d["short"] += int(r[i_short])
        d["wait"] += int(r[i_wait]); d["loc"] += int(r[i_loc]); d["glob"] += int(r[i_glob])
tot = sum(d["s"] for d in lines.values()); tie = sum(d["ie"] for d in lines.values())
print("total samples %d, instructions %d" % (tot, tie))
print("%5s %6s %6s %6s %6s %6s %6s %9s %9s  %s" % ("line", "samp%", "inst%", "long%", "short%", "wait%", "noins%", "locsec", "globsec", "source"))
for (fn, ln), d in sorted(lines.items(), key=lambda x: -x[1]["s"])[:top]:
    print("%5d %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %9.2e %9.2e  %s" % (ln, 100 * d["s"] / tot, 100 * d["ie"] / tie, 100 * d["long"] / tot,
          100 * d["short"] / tot, 100 * d["wait"] / tot, 100 * d["noinst"] / tot, d["loc"], d["glob"], (fn[:10] + ": " if not fn.startswith("mapper_k") else "") + d["src"].strip()[:90]))
print("local sectors by line:")
for (fn, ln), d in sorted(lines.items(), key=lambda x: -x[1]["loc"])[:12]:
    print("%5d %9.2e %s" % (ln, d["loc"], d["src"].strip()[:100]))
print("global sectors by line:")
for (fn, ln), d in sorted(lines.items(), key=lambda x: -x[1]["glob"])[:12]:
    print("%5d %9.2e %s" % (ln, d["glob"], d["src"].strip()[:100]))
