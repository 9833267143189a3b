#!/usr/bin/env python3
"""Summarise ncu exports under gpurun_out/ into tracked files under profiles/.
usage: tools/summarize_ncu.py <round-tag> <log_n of the capture> <capture tags...>"""
import collections, csv, json, os, shutil, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    tag, log_n, caps = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    summary, traffic = [], {}
    for cap in caps:
        path = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.raw.csv" % (tag, cap))
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            ent = collections.OrderedDict(kernel=short, capture_log_n=log_n)
            for k in hdr:
                if k in KEYS or k.split(".")[0] in ("smsp__average_warp_latency_issue_stalled_long_scoreboard",):
                    ent[k] = "%s %s" % (r[idx[k]], units[idx[k]])
            rd = float(r[idx["dram__bytes_read.sum"]]) * SCALE[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * SCALE[units[idx["dram__bytes_write.sum"]]]
            ent["dram_bytes_per_launch"] = rd + wr
            ent["dram_bytes_per_unit"] = (rd + wr) / (1 << log_n)
            summary.append(ent)
            key = short.split("<")[0]
            traffic.setdefault(key, []).append((rd + wr) / (1 << log_n))
        for ext in (".details.txt",):
            src = os.path.join(ROOT, "gpurun_out", "prof_%s_%s%s" % (tag, cap, ext))
            if os.path.exists(src):
                shutil.copy(src, os.path.join(ROOT, "profiles", "%s_%s%s" % (tag, cap, ext)))
    json.dump(summary, open(os.path.join(ROOT, "profiles", "%s_ncu_summary.json" % tag), "w"), indent=1)
    tr = {k: {"bytes_per_unit": sum(v) / len(v), "bytes_per_unit_all_launches": sum(v), "capture_log_n": log_n, "launches": len(v),
              "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch / 2^%d" % log_n}
          for k, v in traffic.items()}
    json.dump(tr, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    for e in summary:
        print(e["kernel"][:50], e.get("gpu__time_duration.sum"), "%.1f B/unit" % e["dram_bytes_per_unit"])


main()
