"""Sums DRAM bytes and durations per kernel family over ONE steady-state step of an ncu launch list
(scripts/gpu_round_check.sh -> launches.csv) and writes profiles/r02_traffic.json (read by bench.py as roofline.traffic).

    python scripts/traffic_from_launches.py gpurun_out/check/launches.csv profiles/r02_traffic.json
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main(src, dst):
    hdr, per = None, {}
    for r in csv.reader(open(src)):
        if r and r[0] == "ID":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        e = per.setdefault(int(d["ID"]), {"name": d["Kernel Name"], "bytes": 0.0, "us": 0.0})
        v = float(d["Metric Value"].replace(",", "")) * UNIT.get(d["Metric Unit"], 1.0)
        if d["Metric Name"].startswith("dram__bytes"):
            e["bytes"] += v
        elif d["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = v
    ids = sorted(per)
    # one step = from one stage-1 re-layout pair (two nchw_to_cl launches) to the next
    starts = [i for i in ids if "nchw_to_cl" in per[i]["name"] and (i - 1 not in per or "nchw_to_cl" not in per[i - 1]["name"])]
    step = [per[i] for i in ids if starts[0] <= i < starts[1]]

    def fam(pred):
        sel = [e for e in step if pred(e["name"])]
        return len(sel), sum(e["bytes"] for e in sel), sum(e["us"] for e in sel)

    k1 = fam(lambda n: "cost_volume_cl_kernel" in n or "nchw_to_cl" in n or "corr_aggregate" in n)
    conv = fam(lambda n: "conv3d_t" in n)
    vis = fam(lambda n: "vis_fused" in n)
    alln = fam(lambda n: True)
    out = {"source": "%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one steady-state step)" % src,
           "k1_launches_per_step": k1[0], "k1_dram_bytes_per_step": k1[1], "k1_dram_bytes_per_launch": k1[1] / max(k1[0], 1),
           "k1_ncu_us_per_step": k1[2], "conv_launches_per_step": conv[0], "conv_dram_bytes_per_step": conv[1],
           "conv_ncu_us_per_step": conv[2], "vis_dram_bytes_per_step": vis[1], "vis_ncu_us_per_step": vis[2],
           "all_dram_bytes_per_step": alln[1], "all_ncu_us_per_step": alln[2], "launches_per_step": alln[0]}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
