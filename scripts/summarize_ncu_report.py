"""Summarise an `ncu --set full` report (.ncu-rep) into a small JSON for profiles/: per launch the duration, DRAM bytes, pipe
utilisation and occupancy. usage: python scripts/summarize_ncu_report.py report.ncu-rep out.json "<command>" "<workload>" ["note"]"""
import csv, json, subprocess, sys
rep, out, cmd, workload = sys.argv[1:5]
note = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
PICK = {
    "duration": "gpu__time_duration.sum", "dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum",
    "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed", "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size", "block": "launch__block_size", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "inst_executed": "smsp__inst_executed.sum",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_realtime_pct": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct", "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "sm_clock_ghz": "sm__cycles_elapsed.max.per_second", "smem_per_block": "launch__shared_mem_per_block_dynamic",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
launches = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
    for k, m in PICK.items():
        if m in hdr:
            i = hdr.index(m)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if k.startswith("dram_bytes") or k == "duration":
                v *= SCALE.get(units[i], 1.0)
            d[k + ("_us" if k == "duration" else "")] = v
    launches.append(d)
res = {"command": cmd, "workload": workload, "launches": launches}
if launches and all("dram_bytes_read" in l for l in launches):
    res["dram_bytes_per_launch"] = sum(l["dram_bytes_read"] + l.get("dram_bytes_write", 0.0) for l in launches) / len(launches)
if note:
    res["note"] = note
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res)[:600])
