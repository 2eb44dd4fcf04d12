"""Turn gpurun_out/ ncu artefacts into the committed summaries under profiles/.

  python profiles/summarize.py <tag> <launches.csv> <prof.ncu-rep> [kernel-regex]
"""
import collections, csv, io, subprocess, sys

tag, launches, rep = sys.argv[1:4]
out = open(f"profiles/{tag}_summary.md", "w")
txt = open(launches).read()
r = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[row["Metric Unit"]]
    k = row["Kernel Name"].split("(")[0]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
out.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
out.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
out.write(f"\n# {tag}: `ncu --set full` capture of the dominant kernel (per launch)\n\n| metric | unit | " +
          " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data) + "\n")
for i, h in enumerate(hdr):
    if h in want or any(h == w for w in want):
        out.write(f"| {h} | {units[i]} | " + " | ".join(d[i][:60] for d in data) + " |\n")
out.close()
print(open(f"profiles/{tag}_summary.md").read())
