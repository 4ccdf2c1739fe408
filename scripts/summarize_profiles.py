"""Turns gpurun_out/{launches.csv, *.ncu-rep} into the small tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [launches.csv] [report.ncu-rep ...]
"""
import collections, csv, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]
out = ROOT / "profiles"
out.mkdir(exist_ok=True)


def short(name):
    m = re.search(r"(gemm_f16_(?:tn|ws)_kernel<[^>]*>|[a-z0-9_]+_kernel(?:<[^>]*>)?)", name)
    return m.group(1) if m else name[:60]


for arg in sys.argv[2:]:
    p = Path(arg)
    if p.suffix == ".csv":
        lines = [l for l in p.open() if not l.startswith("==")]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for row in csv.DictReader(lines):
            v = float(row["Metric Value"].replace(",", ""))
            v *= {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
            k = short(row["Kernel Name"])
            agg[k][0] += 1
            agg[k][1] += v
        tot = sum(v[1] for v in agg.values())
        with (out / f"{tag}_launches.md").open("w") as f:
            f.write(f"# {tag}: kernel launch list (ncu gpu__time_duration.sum, --clock-control none; serialised, cold cache — compare SHARES)\n\n")
            f.write(f"source: `{p.name}`, {sum(v[0] for v in agg.values())} launches, {tot/1e6:.1f} ms of kernel time\n\n")
            f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| `{k}` | {v[0]} | {v[1]/1e6:.2f} | {v[1]/tot*100:.1f}% | {v[1]/v[0]/1e3:.1f} |\n")
        print("wrote", out / f"{tag}_launches.md")
    elif p.suffix == ".ncu-rep":
        raw = subprocess.run(["ncu", "-i", str(p), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
                "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                "sm__cycles_elapsed.max"]
        idx = [i for i, h in enumerate(hdr) if h in want]
        with (out / f"{tag}_{p.stem}_ncu.md").open("w") as f:
            f.write(f"# {tag}: `ncu --set full --clock-control none` of `{p.name}` (selected metrics per captured launch)\n\n")
            for r in rows[2:]:
                f.write(f"## `{short(r[hdr.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
                for i in idx:
                    if hdr[i] != "Kernel Name":
                        f.write(f"| {hdr[i]} | {r[i]} | {units[i]} |\n")
                f.write("\n")
        print("wrote", out / f"{tag}_{p.stem}_ncu.md")
