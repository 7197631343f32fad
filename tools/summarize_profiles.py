#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01            # writes profiles/r01_*.{md,csv,json}
Reads (if present): gpurun_out/launches.csv (gpu__time_duration per launch), prof_*.ncu-rep
(--set full captures, read with `ncu -i ... --page raw --csv`), bench*.json.
"""
import collections, csv, glob, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(PROF, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__cycles_active.avg", "gpc__cycles_elapsed.max",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("f5::", "").replace("void ", "")


def launches():
    p = os.path.join(OUT, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r); idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict(); n = 0
    rows = []
    for row in r:
        v = float(row[idx["Metric Value"]].replace(",", "")); u = row[idx["Metric Unit"]]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        k = short(row[idx["Kernel Name"]])
        rows.append((row[idx["ID"]], k, row[idx["Grid Size"]], row[idx["Block Size"]], f"{v:.3f}"))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; n += 1
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        w = csv.writer(f); w.writerow(["id", "kernel", "grid", "block", "duration_us"]); w.writerows(rows)
    with open(os.path.join(PROF, f"{tag}_launch_shares.md"), "w") as f:
        f.write(f"# {tag}: launch list of `python bench.py --profile-run` under ncu (first {n} launches: precompute + "
                "the first DiT evaluations of one step; cold-cache, serialised — compare SHARES)\n\n"
                "command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --profile-run` "
                "(mel front-end + 2-point sample + Vocos first, then the bench step)\n\n"
                "| kernel | launches | sum µs | avg µs | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {s:.1f} | {s / c:.2f} | {s / tot:.3f} |\n")
    print("launches:", n, "total us", round(tot, 1))


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")]), "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        res.append(d)
    return res


def fulls():
    allr = {}
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        allr[name] = full(rep)
    if not allr:
        return
    json.dump(allr, open(os.path.join(PROF, f"{tag}_ncu_full.json"), "w"), indent=1)
    with open(os.path.join(PROF, f"{tag}_ncu_full.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` captures (see tools/profile.sh)\n\n")
        for name, recs in allr.items():
            for d in recs:
                f.write(f"## {name}: `{d['kernel']}` grid {d['grid']} block {d['block']}\n\n| metric | value |\n|---|---|\n")
                for k in KEYS:
                    if k in d:
                        f.write(f"| {k} | {d[k]} |\n")
                f.write("\n")
    print("full captures:", {k: len(v) for k, v in allr.items()})


def benches():
    for p in sorted(glob.glob(os.path.join(OUT, "bench*.json"))):
        dst = os.path.join(PROF, f"{tag}_{os.path.basename(p)}")
        txt = open(p).read().strip()
        if txt:
            open(dst, "w").write(txt + "\n")


def hbm_table():
    """Achieved DRAM bandwidth of the HBM-bound kernels (prof_hbm capture) against the measured copy peak."""
    rep = os.path.join(OUT, "prof_hbm.ncu-rep")
    if not os.path.exists(rep):
        return
    peak = 6583.5
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        peak = json.load(open(mp)).get("hbm_gbs", peak)
    def num(v):
        m = re.match(r"([0-9.,]+)\s*(\w*)", v or "0")
        x = float(m.group(1).replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}.get(m.group(2), 1)
    agg = collections.OrderedDict()
    for d in full(rep):
        t = num(d.get("gpu__time_duration.sum")); by = num(d.get("dram__bytes_read.sum")) + num(d.get("dram__bytes_write.sum"))
        a = agg.setdefault(d["kernel"] + " grid " + d["grid"], [0, 0.0, 0.0, d.get("lts__t_sector_hit_rate.pct", ""), d.get("sm__warps_active.avg.pct_of_peak_sustained_active", "")])
        a[0] += 1; a[1] += t; a[2] += by
    with open(os.path.join(PROF, f"{tag}_hbm_kernels.md"), "w") as f:
        f.write(f"# {tag}: HBM / FFT kernels under `ncu --set full` (cold cache per replay): DRAM traffic per launch and achieved GB/s "
                f"against the measured copy peak {peak:.0f} GB/s (MEASURED_PEAKS.json)\n\n"
                "| kernel | launches | avg us | DRAM MB / launch | achieved GB/s | of peak | L2 hit % | warps active % |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, (n, t, by, hit, wa) in agg.items():
            gbs = by / t / 1e9 if t > 0 else 0.0
            f.write(f"| `{k}` | {n} | {t / n * 1e6:.2f} | {by / n / 1e6:.3f} | {gbs:.0f} | {gbs / peak:.3f} | {hit.split()[0] if hit else ''} | {wa.split()[0] if wa else ''} |\n")
    print("hbm kernels:", len(agg))


launches(); fulls(); hbm_table(); benches()
