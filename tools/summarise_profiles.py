"""Turns gpurun_out/ ncu outputs into the committed summaries under profiles/.

  python tools/summarise_profiles.py <tag> <launches.csv> <prof.ncu-rep> [steps_in_launch_list]

  profiles/<tag>_launches.md   per-kernel device time of one step from the
                               `ncu --metrics gpu__time_duration.sum --clock-control none` launch list
  profiles/<tag>_kernels.md    `ncu --set full` capture: duration, DRAM bytes, achieved GB/s, % of the
                               measured HBM peak, registers, occupancy, hit rates, top stall reasons
  profiles/<tag>_traffic.json  DRAM bytes per launch per kernel (bench.py's roofline.traffic)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
nsteps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    peak_src = "measured HBM copy bandwidth %.0f GB/s (MEASURED_PEAKS.json)" % peak
except Exception:
    peak = 6650.0        # /opt/skills/guides/B200_PROFILING.md: fallback when the driver's file is absent
    peak_src = "FALLBACK HBM copy bandwidth 6650 GB/s of B200_PROFILING.md (MEASURED_PEAKS.json absent in this container)"
cmdline = os.environ.get("MGP_PROFILE_CMD", "python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    name = name.replace("mgp::", "")
    return name[:60]


rows = []
lines = [l for l in open(launches) if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r["Metric Unit"]]
        rows.append((int(r["ID"]), short(r["Kernel Name"]), v))
# the last `per` launches = whole steps at the end of the run (the bench's phase-timing steps)
names = [n for _, n, _ in rows]
# one step = distance between consecutive k_drift launches
drifts = [i for i, n in enumerate(names) if n in ("k_drift", "k_drift_sd")]
per = drifts[-1] - drifts[-2] if len(drifts) > 1 else len(rows)
last = rows[drifts[-2] + 1: drifts[-1] + 1] if len(drifts) > 1 else rows
agg = collections.OrderedDict()
for _, n, v in last:
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v in agg.values())
with open(os.path.join(ROOT, "profiles", tag + "_launches.md"), "w") as f:
    f.write("# %s: launch list of one COLA step (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n\n" % tag)
    f.write("Command: `" + cmdline + "` (f(R) + screening, 256^3, double grids), "
            "%d launches captured in the whole run, %d in the step shown (a step without a re-sort unless the sort kernels appear). "
            "ncu serialises launches and runs them cold-cache: compare SHARES with bench.py's phases, not absolutes.\n\n" % (len(rows), per))
    f.write("| kernel | launches | device time (us) | share |\n|---|---:|---:|---:|\n")
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (n, c, v, 100 * v / tot))
    f.write("| **total** | %d | %.1f | 100%% |\n" % (sum(c for c, _ in agg.values()), tot))
    # whole-run totals by kernel
    allagg = collections.defaultdict(lambda: [0, 0.0])
    for _, n, v in rows:
        allagg[n][0] += 1
        allagg[n][1] += v
    f.write("\nWhole run (%d launches):\n\n| kernel | launches | device time (us) |\n|---|---:|---:|\n" % len(rows))
    for n, (c, v) in sorted(allagg.items(), key=lambda x: -x[1][1]):
        f.write("| `%s` | %d | %.1f |\n" % (n, c, v))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
col = {h: i for i, h in enumerate(hdr)}


def g(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return float("nan")


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


seen = collections.OrderedDict()
for r in rr[2:]:
    n = short(r[col["Kernel Name"]])
    d = dict(
        us=to_us(g(r, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]]),
        rd=to_bytes(g(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]]),
        wr=to_bytes(g(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]]),
        regs=g(r, "launch__registers_per_thread"), occ=g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        l1=g(r, "l1tex__t_sector_hit_rate.pct"), l2=g(r, "lts__t_sector_hit_rate.pct"),
        issue=g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        dram_pct=g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        grid=r[col["launch__grid_size"]], block=r[col["launch__block_size"]])
    stalls = []
    for h in hdr:
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
        if m and m.group(1) not in ("selected",):
            v = g(r, h)
            if v == v:
                stalls.append((v, m.group(1)))
    d["stalls"] = ", ".join("%s %.1f" % (nm, v) for v, nm in sorted(stalls, reverse=True)[:3])
    seen.setdefault(n, []).append(d)

traffic = {}
with open(os.path.join(ROOT, "profiles", tag + "_kernels.md"), "w") as f:
    f.write("# %s: `ncu --set full --clock-control none --import-source on` capture of the hand-written kernels\n\n" % tag)
    f.write("Same bench command, one step after one warm-up step; one row per kernel (mean over the captured launches). "
            "`GB/s` = (DRAM read + write) / duration; `of peak` = against the " + peak_src + ".  The `.ncu-rep` itself stays in gpurun_out/ (scratch).\n\n")
    f.write("| kernel | n | us | DRAM rd MB | DRAM wr MB | GB/s | of peak | ncu dram % | regs | warps active % | issue active % | L1 hit % | L2 hit % | grid x block | top stalls (warps per issue) |\n")
    f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|\n")
    for n, ds in seen.items():
        k = len(ds)
        m = {key: sum(d[key] for d in ds) / k for key in ("us", "rd", "wr", "regs", "occ", "l1", "l2", "issue", "dram_pct")}
        gbs = (m["rd"] + m["wr"]) / (m["us"] * 1e-6) / 1e9
        traffic[n] = m["rd"] + m["wr"]
        f.write("| `%s` | %d | %.1f | %.1f | %.1f | %.0f | %.2f | %.1f | %d | %.0f | %.0f | %.0f | %.0f | %s x %s | %s |\n" % (
            n, k, m["us"], m["rd"] / 1e6, m["wr"] / 1e6, gbs, gbs / peak, m["dram_pct"], m["regs"], m["occ"], m["issue"], m["l1"], m["l2"],
            ds[0]["grid"], ds[0]["block"], ds[0]["stalls"]))
json.dump(traffic, open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", tag + "_launches.md")).read()[:3000])
print(open(os.path.join(ROOT, "profiles", tag + "_kernels.md")).read())
