"""Summarise `ncu --page raw --csv` exports of --set full captures: one block per kernel with the metrics DESIGN.md /
VERDICT quote (time, issue / FMA / data-pipe utilisation, shared wavefronts, DRAM bytes, stall mix), and optionally
refresh profiles/roofline_traffic.json from a capture of the bench step.
  python profiles/scripts/ncu_summary.py raw1.csv [raw2.csv ...] [--traffic profiles/roofline_traffic.json]"""
import csv, json, sys

KEYS = [("time us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("regs/thread", "launch__registers_per_thread"),
        ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("FMA pipe busy %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("L1/shared data pipe busy %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("warp instructions", "smsp__inst_executed.sum"),
        ("shared wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("  of which LDS", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
        ("  of which LDGSTS", "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum"),
        ("DRAM read MB", "dram__bytes_read.sum"), ("DRAM write MB", "dram__bytes_write.sum"),
        ("L2 hit rate %", "lts__t_sector_hit_rate.pct")]


def rows_of(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield dict(zip(hdr, r)), dict(zip(hdr, units))


def mb(d, u, key):
    v = float(d[key])
    unit = u.get(key, "")
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


traffic_path = None
args = sys.argv[1:]
if "--traffic" in args:
    i = args.index("--traffic")
    traffic_path = args[i + 1]
    del args[i:i + 2]
traffic = {}
for path in args:
    print(f"== {path}")
    for d, u in rows_of(path):
        name = d["Kernel Name"]
        print(f"{name[:110]}")
        for label, key in KEYS:
            if key not in d or d[key] == "":
                continue
            if key.startswith("dram__bytes"):
                print(f"    {label:28s} {mb(d, u, key):12.2f}")
            else:
                print(f"    {label:28s} {d[key]:>12s} {u.get(key, '')}")
        stalls = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v))
                  for k, v in d.items() if "issue_stalled" in k and k.endswith("_per_issue_active.ratio")
                  and "not_issued" not in k and v not in ("", "n/a")]
        stalls = sorted((s for s in stalls if s[1] >= 0.3 and s[0] != "selected"), key=lambda s: -s[1])
        print("    stalls per issue            " + ", ".join(f"{k} {v:.2f}" for k, v in stalls))
        b = int((mb(d, u, "dram__bytes_read.sum") + mb(d, u, "dram__bytes_write.sum")) * 1e6)
        for fam, tag in (("k_adj_own", "interp_adj"), ("k_own_pack", "interp_adj"), ("k_own_fix", "interp_adj"),
                         ("k_fwd_tiled_2d", "interp_fwd")):
            if fam in name:
                traffic[tag] = traffic.get(tag, 0) + b
if traffic_path and traffic:
    traffic["_source"] = ("ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum per launch; interp_adj = sample "
                          "pre-pass + spread [+ fix-up]), cfg2, " + ", ".join(args))
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("wrote", traffic_path, traffic)
