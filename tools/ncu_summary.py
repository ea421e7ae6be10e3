"""Turns the scratch outputs of a GPU round (gpurun_out/) into the tracked evidence under profiles/ (no GPU needed):
raw ncu pages of the hot kernels, the per-kernel launch list, DRAM traffic per launch for bench.py's roofline, the bench
lines and the achieved test errors.        python tools/ncu_summary.py [round-prefix, default r02]"""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__cycles_active.avg", "smsp__inst_executed.sum"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[unit]


summary, traffic = {}, {"source": f"ncu --set full --clock-control none, profiles/{R}_ncu_*.raw.csv (dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                                 "tree / MLP kernels captured with --schedule 1 = one launch per 65,536 searches, N = 5, S = 100)"}
for tag, key in (("tree", "backup_select"), ("net", "net_recurrent"), ("persist", "search_persistent"), ("env", "env_step"),
                 ("x3", "net_x3_recurrent")):
    rep = os.path.join(SRC, f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    with open(os.path.join(DST, f"{R}_ncu_full_{tag}.raw.csv"), "w") as f:
        f.write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        rec = {h: (r[idx[h]], units[idx[h]]) for h in KEEP if h in idx}
        rec["dram_bytes"] = to_bytes(*rec["dram__bytes_read.sum"]) + to_bytes(*rec["dram__bytes_write.sum"])
        rec["us"] = to_us(*rec["gpu__time_duration.sum"])
        out.append(rec)
    summary[tag] = out
    traffic[key] = {"dram_bytes_per_launch": sum(o["dram_bytes"] for o in out) / len(out), "launches_captured": len(out),
                    "duration_us": sum(o["us"] for o in out) / len(out)}
with open(os.path.join(DST, f"{R}_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
with open(os.path.join(DST, f"{R}_ncu_highlights.md"), "w") as f:
    f.write(f"# ncu --set full highlights ({R}; one row per captured launch; raw pages in {R}_ncu_full_*.raw.csv)\n\n")
    for tag, out in summary.items():
        f.write(f"## {tag}: {out[0]['Kernel Name'][0][:110]}\n\n| metric | " + " | ".join(f"launch {i}" for i in range(len(out))) + " |\n|---|" + "---|" * len(out) + "\n")
        for h in KEEP[1:]:
            if h in out[0]:
                f.write(f"| `{h}` | " + " | ".join(f"{o[h][0]} {o[h][1]}" for o in out) + " |\n")
        f.write("| DRAM read + write | " + " | ".join(f"{o['dram_bytes'] / 1e6:.2f} MB" for o in out) + " |\n\n")
# launch list: aggregated per kernel
for lst_name, out_name in (("launches.csv", "launches_step.csv"), ("launches_8192.csv", "launches_step_8192games.csv")):
  lst = os.path.join(SRC, lst_name)
  if os.path.exists(lst):
      rows = [r for r in csv.reader(open(lst)) if len(r) > 10]
      hdr = rows[0]
      i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
      agg = collections.OrderedDict()
      for r in rows[1:]:
          try:
              us = to_us(r[i_val], r[i_unit])
          except (ValueError, KeyError):
              continue
          name = re.sub(r"\(.*", "", r[i_name]).strip()
          a = agg.setdefault(name, [0, 0.0])
          a[0] += 1
          a[1] += us
      tot = sum(a[1] for a in agg.values())
      with open(os.path.join(DST, f"{R}_{out_name}"), "w") as f:
          f.write("kernel,launches,avg_us,total_us,share\n")
          for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
              f.write(f"\"{k}\",{a[0]},{a[1] / a[0]:.3f},{a[1]:.1f},{a[1] / tot:.4f}\n")
# bench lines, sweeps, test metrics
for src, dst in (("bench_default.json", "bench_default_line.json"), ("bench_reference.json", "bench_reference_line.json"),
                 ("bench_2gpu.json", "bench_2gpu_line.json"), ("bench_extras.json", "bench_extras_line.json"),
                 ("bench_persist.json", "bench_persistent_line.json"), ("plain.json", "plain_bench_line.json"),
                 ("plain_persist.json", "plain_persistent_bench_line.json"), ("plain_8192.json", "plain_8192games_bench_line.json"),
                 ("bench_8gpu.json", "bench_8gpu_line.json"), ("persist_sweep.txt", "persist_sweep.txt"),
                 ("group_sweep.txt", "group_sweep.txt"), ("env_sweep.txt", "env_sweep.txt"), ("env_probe.txt", "env_probe.txt"),
                 ("tmem_bench.txt", "microbench_tmem.txt"), ("mma_bench.txt", "microbench_mma.txt"), ("family_alone.txt", "kernel_families_alone.txt"),
                 ("timeline_net_tc.txt", "timeline_net_tc.txt"), ("timeline_tree.txt", "timeline_tree.txt"),
                 ("gantt_4groups.txt", "gantt_4groups.txt"), ("gantt_server.txt", "gantt_server.txt"),
                 ("timeline_net_tc_server.txt", "timeline_net_tc_server.txt"), ("schedules.txt", "schedules.txt"),
                 ("bench_server.json", "bench_server_line.json"), ("x3_probe.txt", "x3_probe.txt"), ("x3_timeline.txt", "x3_timeline.txt"),
                 ("bench_x3.json", "bench_fp32x3_line.json")):
    if os.path.exists(os.path.join(SRC, src)):
        shutil.copyfile(os.path.join(SRC, src), os.path.join(DST, f"{R}_{dst}"))
tm = os.path.join(SRC, "test_metrics.jsonl")
if os.path.exists(tm):
    recs = [json.loads(l) for l in open(tm) if l.strip()]
    with open(os.path.join(DST, f"{R}_net_errors.json"), "w") as f:
        json.dump({"what": "achieved errors recorded by the GPU tests (tests/conftest.py::record_metric), B200", "records": recs}, f, indent=1)
print("\n".join(sorted(x for x in os.listdir(DST) if x.startswith(R))))
