"""profiles/r0N/prof_*_raw.csv (ncu --page raw --csv) -> ncu_full_summary.json: the per-kernel digest bench.py reads for
`roofline.traffic` and the tables of profiles/r0N/README.md quote.  Usage: python scripts/ncu_digest.py profiles/r02"""
import csv
import json
import re
import sys
from pathlib import Path

FIELDS = {
    "time_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pct_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "warp_inst": "smsp__inst_executed.sum",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def digest(path: Path):
    rows = list(csv.reader(path.open()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"kernel": vals[col["Kernel Name"]]}
    for key, metric in FIELDS.items():
        if metric not in col:
            continue
        v = float(vals[col[metric]].replace(",", "") or 0)
        out[key] = v * SCALE.get(units[col[metric]], 1.0)
    return out


def main(folder):
    folder = Path(folder)
    target = folder / "ncu_full_summary.json"
    summary = json.loads(target.read_text()) if target.exists() else {}
    summary.setdefault("_meta", {"workload": "mvsec_dt1", "batch": 32, "corr": "tf32_f16"})
    for f in sorted(folder.glob("prof_*_raw.csv")):
        name = re.sub(r"^prof_|_raw\.csv$", "", f.name)
        try:
            summary[name] = digest(f)
        except (KeyError, ValueError, IndexError) as e:
            print(f"skipped {f.name}: {e}", file=sys.stderr)
    target.write_text(json.dumps(summary, indent=1) + "\n")
    for k, v in summary.items():
        if k != "_meta":
            print(f"{k:44s} {v.get('time_us', 0):8.1f} us  rd {v.get('dram_rd', 0) / 1e6:7.1f} MB  wr {v.get('dram_wr', 0) / 1e6:7.1f} MB  "
                  f"tensor {v.get('tensor_pct_active', 0):5.1f} %  issue {v.get('issue_pct', 0):5.1f} %  inst {v.get('warp_inst', 0) / 1e6:6.2f} M")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02")
