"""Summarise an `ncu -i X.ncu-rep --page raw --csv` dump: one row per launch with the counters DESIGN.md cites, and
(with --json) the per-kernel DRAM traffic per frame that bench.py reports as roofline.traffic.

usage: python tools/ncu_summary.py raw.csv summary.csv [--json out.json --frames N --note "..."]"""
import csv
import json
import re
import sys

KEEP = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        # tensor-core matcher: whichever of these this ncu version exposes
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]


def main():
    raw, out = sys.argv[1], sys.argv[2]
    opts = dict(zip(sys.argv[3::2], sys.argv[4::2]))
    rows = [r for r in csv.reader(l for l in open(raw) if l.startswith('"'))]
    head, units, body = rows[0], rows[1], rows[2:]
    col = {}
    for want in KEEP:
        for i, h in enumerate(head):
            if h == want or h.endswith("." + want):
                col[want] = i
                break
    kn = head.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name []"] + [f"{k} [{units[col[k]]}]" for k in KEEP if k in col])
        for r in body:
            w.writerow([r[kn]] + [r[col[k]] for k in KEEP if k in col])
    if "--json" in opts:
        frames = int(opts.get("--frames", "1"))
        def to_bytes(v, unit):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        def to_ms(v, unit):
            v = float(v.replace(",", ""))
            return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        dram, ms, launches = {}, {}, {}
        for r in body:
            m = re.search(r"(\w+?)_kernel", r[kn])
            name = m.group(1) if m else r[kn]
            name = {"orb_describe": "orb_describe", "desc_or": "desc_or", "finalize": "match_finalize", "match256": "match", "match_tc": "match", "expand_bits": "match_expand"}.get(name, name)
            b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            dram[name] = dram.get(name, 0.0) + b
            ms[name] = ms.get(name, 0.0) + to_ms(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
            launches[name] = launches.get(name, 0) + 1
        total = sum(ms.values())
        json.dump({"source": opts.get("--note", ""), "frames": frames,
                   "dram_bytes_per_frame": {k: round(v / frames) for k, v in dram.items()},
                   "launches_per_step": launches,
                   "ncu_ms_per_step": {k: round(v, 4) for k, v in ms.items()},
                   "ncu_share_of_step": {k: round(v / total, 4) for k, v in ms.items()}}, open(opts["--json"], "w"), indent=1)


if __name__ == "__main__":
    main()
