"""Summarise an Nsight Compute report (.ncu-rep) into the few numbers DESIGN.md / bench.py cite.

    python profiles/summarize_ncu.py gpurun_out/prof_r1.ncu-rep > profiles/r1_kernels.json
    python profiles/summarize_ncu.py gpurun_out/r2_prof_raw.csv  > profiles/r2_kernels_final.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_realtime",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active": "hmma_inst_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_ghz",
}


# every column whose metric name matches one of these is copied as-is (tensor-pipe activity under its various names,
# the warp-state stall breakdown of the WarpStateStats section)
PATTERNS = ("pipe_tensor", "issue_stalled", "smsp__issue_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible")


def main(path):
    if path.endswith(".csv"):      # already exported with `ncu -i <rep> --page raw --csv` (done on the GPU box to keep gpurun_out small)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    while rows and "ID" not in rows[0]:
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for h, i in list(col.items()):  # some metrics carry a section prefix ("TPC.TriageCompute.<metric>")
        for k in KEYS:
            if h.endswith("." + k) and k not in col:
                col[k] = i
    res = []
    for r in rows[2:]:
        d = {"id": r[col["ID"]], "kernel": r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")}
        for k, name in KEYS.items():
            if k in col:
                v = r[col[k]].replace(",", "")
                try:
                    v = float(v)
                except ValueError:
                    pass
                u = units[col[k]]
                if name.endswith("_MB") and isinstance(v, float):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if name == "time_us" and isinstance(v, float):
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                d[name] = v
        extra = {}
        for h, i in col.items():
            if any(pt in h for pt in PATTERNS) and h not in KEYS:
                v = r[i].replace(",", "")
                try:
                    v = float(v)
                except ValueError:
                    continue
                if v != 0.0:
                    extra[h.split(".", 2)[-1] if h.count(".") > 3 else h] = round(v, 4)
        d["counters"] = extra
        res.append(d)
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
