"""Parses an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k4_hist_build_ws`
log of one boosting iteration into profiles/r01_k4_dram_traffic_<rows>x<F>.json (per-launch average, as bench.py's roofline.traffic)."""
import csv
import json
import sys


def main(path, rows, feats, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = {}
    for r in csv.DictReader(lines):
        if "k4_hist_build_ws" not in r.get("Kernel Name", ""):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        per.setdefault(r["ID"], {})[r["Metric Name"]] = v * scale
    launches = list(per.values())
    rd = sum(x.get("dram__bytes_read.sum", 0.0) for x in launches)
    wr = sum(x.get("dram__bytes_write.sum", 0.0) for x in launches)
    ns = sum(x.get("gpu__time_duration.sum", 0.0) for x in launches)
    d = {"rows": rows, "features": feats, "launches": len(launches), "dram_read_bytes_total": rd, "dram_write_bytes_total": wr,
         "bytes_per_launch_avg": (rd + wr) / max(len(launches), 1), "k4_time_ms_under_ncu": ns / 1e6,
         "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one boosting iteration "
                   "(all K4 launches of one tree) of bench.py --ingest device at this shape"}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
