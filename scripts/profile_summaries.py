"""Regenerates the judged summaries under profiles/ from raw ncu output:
  launches:  python scripts/profile_summaries.py launches LAUNCHES.csv "COMMAND" > profiles/launches_rN.md
  kernel:    python scripts/profile_summaries.py kernel REPORT.ncu-rep "COMMAND" > profiles/ncu_trace_rN.md
             (also prints dram bytes as JSON on stderr for profiles/trace_dram_traffic.json)"""
import csv, io, json, subprocess, sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]


def launches(path, command):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    seq = []
    for r in rows[1:]:
        if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("m3d::<unnamed>::", "").replace("m3d::", "")
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        seq.append((name, v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v))
    print("# ncu launch list: `%s`" % command)
    print("# (cold-cache, serialised: compare shares, not absolutes). Raw CSV: profiles/%s" % path.split("/")[-1])
    # the headline step ends where the appended path-tracing measurements start
    cut = next((i for i, (n, _) in enumerate(seq) if n.startswith("path_") or n.startswith("bidir_")), len(seq))
    for title, part in (("Headline part of the step (C2 first-hit batches: counters pass, warm-up, timed steps, "
                         "host-buffer e2e calls)", seq[:cut]), ("Whole command (headline + appended path-tracing "
                                                                "measurements, up to the -c limit)", seq)):
        agg = OrderedDict()
        for name, ms in part:
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ms
        tot = sum(a[1] for a in agg.values())
        print("\n## %s\n\n| kernel | launches | total ms | share |\n|---|---|---|---|" % title)
        for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("| `%s` | %d | %.3f | %.1f%% |" % (name[:80], c, ms, 100 * ms / tot))


def kernel(path, command):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full summary (profiles/%s): `%s`\n" % (path.split("/")[-1], command))
    traffic = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0]
        print("## %s\n| metric | unit | value |\n|---|---|---|" % name)
        for m in METRICS:
            if m in ix:
                print("| %s | %s | %s |" % (m, units[ix[m]], r[ix[m]]))
        print()

        def val(m):
            v, u = float(r[ix[m]].replace(",", "")), units[ix[m]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        traffic[name] = {"dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                         "duration": r[ix["gpu__time_duration.sum"]] + " " + units[ix["gpu__time_duration.sum"]]}
    print(json.dumps(traffic, indent=1), file=sys.stderr)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
