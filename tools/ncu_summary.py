"""ncu `--page raw --csv` export -> the per-kernel columns DESIGN.md / profiles/ quote.

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv [skip_first_n] > summary.json

`skip_first_n` drops the warm-up launches of a two-pass capture (tools/profile_kernels.py prints the launch count)."""
import csv
import json
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__grid_size"]
# everything is reported in ms / GB whatever unit ncu picked for the column
SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0,
         "Tbyte": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    out = []
    for r in rows[hi + 2:]:
        if not r or not r[0].isdigit() or int(r[0]) < skip:
            continue
        k = {"id": int(r[0]), "kernel": r[hdr.index("Kernel Name")]}
        for c in COLS:
            if c not in hdr:
                k[c] = None
                continue
            i = hdr.index(c)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                k[c] = None
                continue
            k[c] = v * SCALE.get(units[i], 1.0)
        rd, wr = k.get("dram__bytes_read.sum") or 0.0, k.get("dram__bytes_write.sum") or 0.0
        k["dram_GB"] = round(rd + wr, 3)
        t = k.get("gpu__time_duration.sum")
        k["dram_TBps"] = round((rd + wr) / t, 3) if t else None   # GB / ms = TB/s
        out.append(k)
    json.dump({"source": "ncu --set full --clock-control none; " + sys.argv[1], "units": {"time": "ms", "bytes": "GB"},
               "kernels": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
