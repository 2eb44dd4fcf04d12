"""Per-launch table of an `ncu --set full` capture that holds several different kernels.

  python profiles/summarize_misc.py <tag> <launches.csv> <prof.ncu-rep>
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
    agg[row["Kernel Name"].split("(")[0]][0] += 1; agg[row["Kernel Name"].split("(")[0]][1] += v
tot = sum(v[1] for v in agg.values())
out.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
out.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
out.write(f"\n# {tag}: `ncu --set full` captures (one row per launch; units as ncu reports them)\n\n| kernel | " +
          " | ".join(f"{n} [{units[col[m]]}]" for m, n in want if m in col) + " |\n|---|" + "---|" * sum(m in col for m, _ in want) + "\n")
seen = collections.Counter()
for d in data:
    name = d[col["Kernel Name"]].split("(")[0]
    seen[name] += 1
    if seen[name] > 3:
        continue
    out.write(f"| `{name}` | " + " | ".join(d[col[m]][:12] for m, _ in want if m in col) + " |\n")
out.close()
print(open(f"profiles/{tag}_summary.md").read())
