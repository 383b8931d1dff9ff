"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
usage: python scripts/ncu_summary.py <tag> [launches.csv] [name=report.ncu-rep ...]"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__inst_executed.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__inst_executed.sum",
        "launch__occupancy_limit_shared_mem"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        a = agg.setdefault(row["Kernel Name"].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.write("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n")
    out.write("%-72s %6s %12s %7s\n" % ("kernel", "n", "total_us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.write("%-72s %6d %12.1f %6.1f%%\n" % (k, a[0], a[1], 100 * a[1] / tot))
    out.write("%-72s %6s %12.1f\n\n" % ("TOTAL", "", tot))


def report(name, path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        out.write("# %s: could not read %s\n" % (name, path))
        return
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name
        out.write("# %s — `ncu --set full --clock-control none` — %s\n" % (name, kn[:90]))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.write("%-78s %18s %s\n" % (w, vals[i], units[i]))
        out.write("\n")


def traffic(paths, tag):
    """profiles/kernel_traffic.json: dram bytes (read + write) per launch of the kernels bench.py reports a roofline for."""
    import json
    out = {}
    for path in paths:
        traffic_one(path, tag, out)
    json.dump(out, open("profiles/kernel_traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


def traffic_one(path, tag, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    names = {"ppm_gather_kernel": "ppm.gather", "ppm_direct_output_kernel": "ppm.direct_output"}
    seen_trace = 0
    tunit = rows[1][hdr.index("gpu__time_duration.sum")]
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tunit, 1.0)
    for vals in rows[2:]:
        kn = vals[hdr.index("Kernel Name")].split("(")[0]
        key = names.get(kn)
        if "trace_shadow_queue_kernel" in kn:
            key = "trace.shadow"
        if "trace_closest_kernel" in kn:
            # the node format is a template argument: <0> = fp32 nodes (primary rays of an iteration), <1> = quantised nodes (photon bounces);
            # the first <1> launch after a <0> launch is photon bounce 0 of the same iteration
            if "<0>" in kn or "(bool)0" in kn:
                key, seen_trace = "trace.camera0", 1
            elif seen_trace == 1:
                key, seen_trace = "trace.photon0", 2
        if not key or key in out:
            continue
        rd, wr = float(vals[hdr.index("dram__bytes_read.sum")]), float(vals[hdr.index("dram__bytes_write.sum")])
        ur, uw = rows[1][hdr.index("dram__bytes_read.sum")], rows[1][hdr.index("dram__bytes_write.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        def metric(name):
            return float(vals[hdr.index(name)]) if name in hdr else None
        out[key] = {"dram_bytes": rd * scale[ur] + wr * scale[uw], "gpu_time_ms_under_ncu": float(vals[hdr.index("gpu__time_duration.sum")]) * tscale,
                    "issue_active_pct": metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "l1_data_pipe_wavefronts_pct": metric("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                    "l2_throughput_pct": metric("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "dram_throughput_pct": metric("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "lanes_per_instruction": metric("smsp__thread_inst_executed_per_inst_executed.ratio"),
                    "source": "profiles/%s (ncu --set full --clock-control none, one launch)" % tag}


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1].startswith("-"):      # `--help` used to be taken for a tag and wrote profiles/--help_ncu_summary.txt
        print(__doc__ + "\n       python scripts/ncu_summary.py traffic <summary file name> report.ncu-rep [...]")
        sys.exit(0 if len(sys.argv) > 1 and sys.argv[1] in ("-h", "--help") else 2)
    if sys.argv[1] == "traffic":
        traffic(sys.argv[3:], sys.argv[2])     # traffic <summary file name> report.ncu-rep [...]
        sys.exit(0)
    tag = sys.argv[1]
    with open("profiles/%s_ncu_summary.txt" % tag, "w") as out:
        for a in sys.argv[2:]:
            if "=" in a:
                n, p = a.split("=", 1)
                report(n, p, out)
            else:
                launches(a, out)
    print(open("profiles/%s_ncu_summary.txt" % tag).read())
