#!/usr/bin/env python
"""profiles/kernel_traffic.json from `ncu --set full` captures of the bench command: DRAM bytes per launch of K1 and K2.

    python tools/ncu_traffic.py <workload>=<report.ncu-rep>:<columns> ...

``columns`` = snapshots (+ halo) in the captured launch; ``bench.py`` only quotes the traffic when its own launches have
the same shape.  Units are normalised to bytes and microseconds.
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3,
         "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def main():
    out_path = ROOT / "profiles" / "kernel_traffic.json"
    table = json.loads(out_path.read_text()) if out_path.exists() else {}
    for spec in sys.argv[1:]:
        wl, rest = spec.split("=", 1)
        rep, cols = rest.rsplit(":", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]

        def val(r, key):
            i = hdr.index(key)
            return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)

        entry = {}
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            short = "k1_stage" if "k1_stage" in name else "k2_wall" if "k2_wall" in name else None
            if short is None or short in entry:
                continue
            rd, wr, us = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
            entry[short] = {"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                            "duration_us_under_ncu": round(us, 3), "columns": int(cols),
                            "dram_gbs_under_ncu": round((rd + wr) / us / 1e3, 1),
                            "fp64_pipe_active_pct": round(val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 1),
                            "kernel": name.split("(")[0].replace("<unnamed>::", "").replace("void ", ""),
                            "source": f"profiles/{Path(rep).stem}.md (ncu --set full --clock-control none, cold-cache "
                                      "replay of the bench command)"}
        table[wl] = entry
    out_path.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
    print(json.dumps(table, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
