"""Per-launch table of an `ncu --set full` capture that holds several different kernels.

  python profiles/summarize_misc.py <tag> <launches.csv> <prof.ncu-rep> [more .ncu-rep ...]
"""
import collections, csv, io, subprocess, sys

tag, launches = sys.argv[1:3]
reps = sys.argv[3:]
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
hdr, data = None, []                      # data rows carry "value unit" cells: ncu scales the units per report
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    if hdr is None:
        hdr = rows[0]
    idx = {h: i for i, h in enumerate(rows[0])}
    u = rows[1]
    for r in rows[2:]:
        data.append([(r[idx[h]][:12] + (" " + u[idx[h]] if u[idx[h]] and h != "Kernel Name" else "")) if h in idx else "" for h in hdr]
                    if True else None)
        data[-1][hdr.index("Kernel Name")] = r[idx["Kernel Name"]]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"), ("lts__t_sector_hit_rate.pct", "L2 hit")]
out.write(f"\n# {tag}: `ncu --set full` captures (one row per launch; each cell with the unit ncu reported)\n\n| kernel | " +
          " | ".join(n for m, n in want if m in col) + " |\n|---|" + "---|" * sum(m in col for m, _ in want) + "\n")
seen = collections.Counter()
for d in data:
    name = d[col["Kernel Name"]].split("(")[0]
    seen[name] += 1
    if seen[name] > 2:
        continue
    out.write(f"| `{name}` | " + " | ".join(d[col[m]].replace("register/thread", "").strip() for m, _ in want if m in col) + " |\n")
out.close()
print(open(f"profiles/{tag}_summary.md").read())
