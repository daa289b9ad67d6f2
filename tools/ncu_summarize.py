"""Turns the raw ncu outputs of tools/run_measure.sh into the tracked summaries under profiles/:

    python tools/ncu_summarize.py TAG STEPS_IN_LAUNCH_LIST [ROUND]

  gpurun_out/launches_TAG.csv -> profiles/rNN_launches_16m.csv     (per kernel: launches, mean us, per step, share)
  gpurun_out/raw_TAG.csv      -> profiles/rNN_ncu_top_kernels_16m.csv (selected metrics of the two sweeps)
                              -> profiles/kernel_metrics.json (what bench.py's roofline block reads: DRAM bytes per
                                 launch, FP32-pipe / LSU / issue utilisation and the captured duration of each sweep)
"""
import collections, csv, json, re, sys
tag, steps = sys.argv[1], int(sys.argv[2])
RND = sys.argv[3] if len(sys.argv) > 3 else "r02"
rows = [r for r in csv.reader(l for l in open("gpurun_out/launches_%s.csv" % tag) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
SETUP = ("k_iota", "k_pack_state", "k_unpack_xyz", "k_pack_vel", "k_pack_pos", "k_gather_vel")        # upload / e2e leg, not part of a device-resident step
t = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    if name.startswith("cub::"):
        name = name.split("<")[0]
    if name.startswith(SETUP):
        continue
    t.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1e3)
# the bench runs warm-up + timed + timer + e2e steps; every one of them launches each step kernel once
n_steps = max(len(v) for v in t.values())
total = sum(sum(v) / n_steps for v in t.values())
with open("profiles/%s_launches_16m.csv" % RND, "w") as f:
    f.write("# ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` over\n"
            "# `python bench.py --steps %d --warmup 3 --no-cpu-baseline` (16.7M dam-break, 1 B200), tag %s.  Times are\n"
            "# cold-cache / serialised: compare SHARES.  mean_us = mean per launch; per_step = launches per step x mean.\n" % (steps, tag))
    f.write("kernel,launches,mean_us,per_step_us,share_of_step\n")
    for k, v in t.items():
        per = sum(v) / n_steps
        f.write("%s,%d,%.1f,%.1f,%.3f\n" % (k, len(v), sum(v) / len(v), per, per / total))
    f.write("TOTAL,,,%.1f,1.000\n" % total)
raw = list(csv.reader(open("gpurun_out/raw_%s.csv" % tag)))
h, u, data = raw[0], raw[1], raw[2:]
keep = [m for m in h if re.match(r"(gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|"
                                 r"l1tex__data_pipe_lsu_wavefronts.avg.pct|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|"
                                 r"lts__throughput.avg.pct|launch__(block_size|grid_size|registers_per_thread$|occupancy_limit)|"
                                 r"sm__inst_executed_pipe_(alu|fma|lsu).avg.pct_of_peak_sustained_active|"
                                 r"sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active|"
                                 r"sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed|"
                                 r"sm__throughput.avg.pct|sm__warps_active.avg.pct|smsp__inst_executed.sum$|"
                                 r"smsp__issue_active.avg.pct|smsp__thread_inst_executed_per_inst_executed.ratio|"
                                 r"smsp__average_warps_issue_stalled_(barrier|long_scoreboard|math_pipe_throttle|"
                                 r"no_instruction|not_selected|short_scoreboard|wait)_per_issue_active)", m)]
with open("profiles/%s_ncu_top_kernels_16m.csv" % RND, "w") as f:
    f.write("# ncu --set full --clock-control none, 16.7M dam-break (bench.py workload), 1 B200, tag %s\n" % tag)
    f.write("metric,unit," + ",".join(re.sub(r"\(.*", "", d[h.index("Kernel Name")]).replace("void <unnamed>::", "") for d in data) + "\n")
    for m in keep:
        i = h.index(m)
        f.write("%s,%s,%s\n" % (m, u[i], ",".join(d[i] for d in data)))

def val(d, name):
    try:
        return float(d[h.index(name)].replace(",", ""))
    except (ValueError, IndexError):
        return None


def to_bytes(d, name):
    i = h.index(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[i]]
    return float(d[i].replace(",", "")) * scale


kernels = {}
for d in data:
    kname = d[h.index("Kernel Name")]
    key = "density" if "k_density" in kname else "force" if "k_force" in kname else None
    if not key:
        continue
    tunit = u[h.index("gpu__time_duration.sum")]
    t = val(d, "gpu__time_duration.sum") * {"us": 1e-3, "ms": 1.0, "ns": 1e-6}[tunit]
    kernels[key] = {
        "kernel": re.sub(r"\(.*", "", kname).replace("void <unnamed>::", ""),
        "dram_bytes": to_bytes(d, "dram__bytes_read.sum") + to_bytes(d, "dram__bytes_write.sum"),
        "ncu_ms": t,
        "inst_executed": val(d, "smsp__inst_executed.sum"),
        "fma_pipe_pct": val(d, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "lsu_wavefront_pct": val(d, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "issue_pct": val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_pct": val(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    }
with open("profiles/kernel_metrics.json", "w") as f:
    json.dump({"particles": 16777216, "nu": 40.0, "source": "profiles/%s_ncu_top_kernels_16m.csv, tag %s" % (RND, tag),
               "kernels": kernels}, f, indent=1)
print(open("profiles/%s_launches_16m.csv" % RND).read())
print(json.dumps(kernels, indent=1))
