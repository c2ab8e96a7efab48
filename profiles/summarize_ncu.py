"""Summarises `ncu --page raw --csv` exports (gpurun_out/prof_r02*_raw.csv) into profiles/ncu_r02_top_kernels.csv and profiles/traffic.json.
    python profiles/summarize_ncu.py gpurun_out/prof_r02_raw.csv [more.csv ...]"""
import csv
import json
import os
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("smsp__issue_active.avg.pct", "issue_active_pct"), ("launch__occupancy_limit_registers", "occ_limit_regs"),
        ("launch__occupancy_limit_shared_mem", "occ_limit_smem"), ("launch__block_size", "block"), ("launch__grid_size", "grid")]

HERE = os.path.dirname(os.path.abspath(__file__))
rows_out, traffic = [], {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0].replace("void ", "")
        rec = {"kernel": short, "source": os.path.basename(path)}
        for metric, key in WANT:
            if metric in hdr:
                i = hdr.index(metric)
                rec[key] = r[i]
                rec[key + "_unit"] = units[i]
        rows_out.append(rec)

def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

def to_us(v, unit):
    v = float(v)
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)

with open(os.path.join(HERE, "ncu_r02_top_kernels.csv"), "w") as f:
    f.write("kernel,time_us,dram_read_GB,dram_write_GB,regs,warps_active_pct,dram_pct,fp64_pipe_pct,issue_active_pct,block,grid,source\n")
    for r in rows_out:
        rd = to_bytes(r["dram_read"], r["dram_read_unit"]) / 1e9
        wr = to_bytes(r["dram_write"], r["dram_write_unit"]) / 1e9
        f.write("%s,%.1f,%.4f,%.4f,%s,%s,%s,%s,%s,%s,%s,%s\n" % (r["kernel"].replace(",", ";"), to_us(r["time"], r["time_unit"]), rd, wr, r.get("regs"),
                                                          r.get("warps_active_pct"), r.get("dram_pct"), r.get("fp64_pipe_pct"),
                                                          r.get("issue_active_pct"), r.get("block"), r.get("grid"), r["source"]))
        traffic.setdefault(r["kernel"], []).append(rd + wr)
print(open(os.path.join(HERE, "ncu_r02_top_kernels.csv")).read())
