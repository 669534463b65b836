"""Turn `ncu --set full` reports into the committed evidence: a compact CSV of the roofline-relevant metrics per captured
launch, and profiles/ncu_traffic.json (DRAM bytes per launch of the dominant kernel, which bench.py reports as
roofline.traffic instead of a literal).

usage (here, no GPU needed):  python profiles/summarize_ncu.py profiles/r2/ncu_metrics.csv gpurun_out/<tag>/*.ncu-rep
The raw page of each report is read with `ncu -i <rep> --page raw --csv`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def read(rep):
    if rep.endswith(".csv.gz"):          # a raw-page export made on the GPU box (the report itself was too big to bring back)
        import gzip
        raw = gzip.open(rep, "rt").read()
    elif rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {"report": os.path.basename(rep), "kernel": r[col["Kernel Name"]].split("(")[0], "grid": r[col["Grid Size"]],
             "block": r[col["Block Size"]]}
        for m in METRICS:
            if m not in col:
                continue
            v, u = r[col[m]].replace(",", ""), units[col[m]]
            try:
                v = float(v)
            except ValueError:
                continue
            if m.startswith("dram__bytes"):
                v *= TO_BYTES.get(u, 1.0)
            elif m == "gpu__time_duration.sum":
                v *= TO_US.get(u, 1.0)
                m = "gpu__time_duration_us"
            d[m] = v
        out.append(d)
    return out


def main(dst, reps):
    rows = [r for rep in reps for r in read(rep)]
    keys = ["report", "kernel", "grid", "block"] + sorted({k for r in rows for k in r} - {"report", "kernel", "grid", "block"})
    with open(dst, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(rows)
    # the dominant kernel family: the largest forward launch captured
    fwd = [r for r in rows if "conv_tma_fwd_kernel" in r["kernel"] and "dram__bytes_read.sum" in r]
    if fwd:
        r = max(fwd, key=lambda r: r.get("gpu__time_duration_us", 0.0))
        traffic = {"conv_tma_fwd": {
            "dram_bytes_per_launch": r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"],
            "dram_read": r["dram__bytes_read.sum"], "dram_write": r["dram__bytes_write.sum"],
            "kernel": r["kernel"], "grid": r["grid"], "duration_us_under_ncu": r.get("gpu__time_duration_us"),
            "tensor_pipe_active_pct": r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "source": f"{os.path.relpath(dst, ROOT)} ({r['report']}: ncu --set full --clock-control none of "
                      "profiles/conv_microbench.py --only s9_cell1_h_ur, the per-timestep h-half update|reset GEMM of "
                      "the 32x32 ConvGRU stage at B=64: M=65536, Cin=256, Cout=512, 5x5)"}}
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    for r in rows:
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
