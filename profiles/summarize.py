#!/usr/bin/env python
"""Summarise an ncu report into profiles/<tag>_summary.md (+ dram_traffic.json / kernel_counters.json consumed by bench.py).
usage: python profiles/summarize.py gpurun_out/prof_<tag>.ncu-rep <tag> [launches.csv] [--config C] [--gpus N]
dram_traffic.json is keyed by the capture's workload, {"cfg<C>_n<N>": {kernel: bytes per launch}}: bench.py quotes `roofline.traffic`
only on a line of the same config at the same GPU count.  kernel_counters.json holds the config-3, one-GPU counters."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "global load wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU (conversions, MUFU) %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
]


def main():
    argv = list(sys.argv)
    cfg, gpus = 3, 1
    for flag in ("--config", "--gpus"):
        if flag in argv:
            i = argv.index(flag)
            if flag == "--config":
                cfg = int(argv[i + 1])
            else:
                gpus = int(argv[i + 1])
            del argv[i:i + 2]
    sys.argv = argv
    if "--traffic-csv" in argv:      # a `--metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list: DRAM bytes of each kernel's first launch
        path = argv[argv.index("--traffic-csv") + 1]
        per = {}
        seen = {}
        for row in csv.DictReader(l for l in open(path) if not l.startswith("==")):
            if row.get("Metric Name") not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                continue
            k = row["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
            if seen.setdefault(k, row["ID"]) != row["ID"]:
                continue
            v = float(row["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row["Metric Unit"], 1)
            per[k] = per.get(k, 0.0) + v
        here = os.path.dirname(os.path.abspath(__file__))
        tp = os.path.join(here, "dram_traffic.json")
        try:
            all_traffic = json.load(open(tp))
        except Exception:
            all_traffic = {}
        all_traffic[f"cfg{cfg}_n{gpus}"] = per
        json.dump(all_traffic, open(tp, "w"), indent=1, sort_keys=True)
        print(json.dumps(per))
        return
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu summary `{tag}` ({os.path.basename(rep)}; `ncu --set full --clock-control none`)\n"]
    traffic = {}
    counters = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        out.append(f"\n## {name}\n\n| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in idx:
                out.append(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        try:
            def b(k):
                v, u = float(r[idx[k]]), units[idx[k]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic[name.split("<")[0]] = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
        except Exception:
            pass
        try:    # warp instructions per launch: bench.py turns them into a live fraction of the SM issue rate
            counters.setdefault(name.split("<")[0], {})
            if not counters[name.split("<")[0]]:
                counters[name.split("<")[0]] = {"warp_instructions": float(r[idx["smsp__inst_executed.sum"]]),
                                            "active_threads_per_inst": float(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                                            "issue_slots_busy_pct_ncu": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]])}
        except Exception:
            pass
    here = os.path.dirname(os.path.abspath(__file__))
    if len(sys.argv) > 3 and os.path.exists(sys.argv[3]):
        out.append("\n## launch list (gpu__time_duration.sum per launch, cold-cache, serialised)\n")
        agg = {}
        for row in csv.DictReader(l for l in open(sys.argv[3]) if not l.startswith("==")):
            if row.get("Metric Name") == "gpu__time_duration.sum":
                k = row["Kernel Name"].split("(")[0]
                v = float(row["Metric Value"].replace(",", ""))
                u = row["Metric Unit"]
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
                a = agg.setdefault(k, [0, 0.0])
                a[0] += 1; a[1] += v
        tot = sum(a[1] for a in agg.values())
        out.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| {k} | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% |")
    open(os.path.join(here, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
    tp = os.path.join(here, "dram_traffic.json")
    try:
        all_traffic = json.load(open(tp))
        if not all(isinstance(v, dict) for v in all_traffic.values()):
            all_traffic = {}
    except Exception:
        all_traffic = {}
    all_traffic[f"cfg{cfg}_n{gpus}"] = traffic
    json.dump(all_traffic, open(tp, "w"), indent=1, sort_keys=True)
    if cfg == 3 and gpus == 1:
        json.dump(counters, open(os.path.join(here, "kernel_counters.json"), "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
