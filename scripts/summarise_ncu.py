"""Summarise ncu --set full captures (gpurun_out/*.ncu-rep) into small tracked files under profiles/.

    python scripts/summarise_ncu.py gpurun_out/prof_k1_r1f.ncu-rep profiles/r01_k1_full.csv [family-key]

Writes one CSV row per captured launch with the metrics the roofline discussion uses, and (when a
family key is given) records the mean DRAM traffic per launch in profiles/r01_traffic.json, which
bench.py reports as `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    family = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), short, units[hdr.index(m)]) for m, short in METRICS if m in hdr]
    traffic = []
    stall_cols = [(i, h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
    with open(out, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow([short + ("_us" if short == "time" else "_bytes" if short.startswith("dram_") and not short.endswith("pct") else "") for _, short, _ in cols] + ["top_stalls(pct of warp-active cycles)"])
        for r in rows[2:]:
            vals = {}
            line = []
            for i, short, unit in cols:
                v = r[i]
                if short in ("time", "dram_read", "dram_write"):
                    v = float(v.replace(",", "")) * UNIT_SCALE.get(unit, 1.0)
                    vals[short] = v
                    v = "%.1f" % v
                elif short == "kernel":
                    v = v.split("(")[0].replace("void ", "")[:90]
                line.append(v)
            stalls = []
            for i, h in stall_cols:
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_warp_active.pct", "")))
                except ValueError:
                    pass
            stalls.sort(reverse=True)
            line.append("; ".join("%s %.0f" % (n, v) for v, n in stalls[:6]))
            wr.writerow(line)
            traffic.append(vals.get("dram_read", 0.0) + vals.get("dram_write", 0.0))
    if family and traffic:
        path = os.path.join(os.path.dirname(out), "r01_traffic.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[family] = sum(traffic) / len(traffic)
        json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
