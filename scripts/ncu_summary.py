"""Turn ncu output into the markdown tables kept under profiles/.

  python scripts/ncu_summary.py launches <launches.csv> <out.md> "<command line>"
  python scripts/ncu_summary.py full <report.ncu-rep> [<report2.ncu-rep> ...] > section.md
  python scripts/ncu_summary.py raw <report.raw.csv[.gz]> ... > section.md     (csv made on the GPU box by ncu_export.sh)

`launches`: csv log of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`.
`full`: reports of `ncu --set full --clock-control none --import-source on -k regex:... -c N`; read with
`ncu -i X --page raw --csv` (needs ncu on PATH; no GPU)."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe % (inst executed, of peak)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe % (cycles active)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def launches(path, out, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
              "second": 1e6, "s": 1e6}.get(r[ui], 1.0)                      # -> microseconds
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# launch list - `%s`\n\n" % cmd)
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live\n"
                "`stage_ms_per_step`, not absolutes.  `cutlass ... d884gemm` rows are bench.py's torch.matmul FP64 peak\n"
                "probe, not the engine.  Raw csv next to this file.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k[:100], n, us, 100 * us / tot))


def full(reports, from_csv=False):
    for rep in reports:
        if from_csv:
            import gzip
            raw = (gzip.open(rep, "rt") if rep.endswith(".gz") else open(rep)).read()
        else:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        for d in data:
            print("### `%s`  (%s)\n" % (d[hdr.index("Kernel Name")][:110], rep.split("/")[-1]))
            print("| metric | value |\n|---|---|")
            for key, label in METRICS:
                if key in hdr:
                    i = hdr.index(key)
                    print("| %s (`%s`) | %s %s |" % (label, key, d[i], units[i]))
            top = sorted(((float(d[hdr.index(k)].replace(",", "")), k) for k in stall), reverse=True)[:5]
            print("| top warp stalls (warps per issue-active cycle) | %s |" % ", ".join(
                "%s %.2f" % (k.split("stalled_")[1].split("_per")[0], v) for v, k in top))
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "raw":
        full(sys.argv[2:], from_csv=True)
    else:
        full(sys.argv[2:])
