"""Turn the .ncu-rep captures brought back in gpurun_out/ into the tracked evidence under profiles/:
  python scripts/ncu_extract.py <tag> <rep> [<rep> ...]
For every report: profiles/<tag>_<name>_raw.csv = `ncu -i <rep> --page raw --csv` restricted to the metrics the roofline
figures are computed from; and profiles/r2_ncu_metrics.json (read by bench.py) is updated with the per-kernel numbers."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__average_warp_latency_issue_stalled_barrier.pct"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    return hdr, units, rows[2:]


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    mpath = os.path.join(ROOT, "profiles", "r2_ncu_metrics.json")
    M = json.load(open(mpath)) if os.path.exists(mpath) else {}
    for rep in reps:
        hdr, units, rows = raw_rows(rep)
        cols = [i for i, h in enumerate(hdr) if h in KEEP or any(h.startswith(k) for k in ("sm__pipe_tensor", "sm__pipe_fp64", "dram__bytes"))]
        name = os.path.splitext(os.path.basename(rep))[0]
        with open(os.path.join(ROOT, "profiles", "%s_%s_raw.csv" % (tag, name)), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow([hdr[i] for i in cols]); w.writerow([units[i] for i in cols])
            for r in rows:
                w.writerow([r[i] for i in cols])
        H = {h: i for i, h in enumerate(hdr)}

        def val(r, key):
            if key not in H:
                return None
            try:
                v = float(r[H[key]].replace(",", ""))
            except ValueError:
                return None
            u = units[H[key]]
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
            return v * scale
        for r in rows:
            kn = r[H["Kernel Name"]]
            rec = {"duration_s": val(r, "gpu__time_duration.sum"), "dram_read_bytes": val(r, "dram__bytes_read.sum"),
                   "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                   "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   "fp64_pipe_pct": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                   "regs": val(r, "launch__registers_per_thread"), "grid": val(r, "launch__grid_size")}
            M.setdefault("kernels", {}).setdefault(name, []).append({"kernel": kn[:80], **rec})
            print(name, kn[:60], rec)
    json.dump(M, open(mpath, "w"), indent=1)


if __name__ == "__main__":
    main()
