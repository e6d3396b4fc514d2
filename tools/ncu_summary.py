"""Summarise `ncu --set full` reports into a markdown table (run HERE, on the CPU box: ncu reads .ncu-rep files offline).
    python tools/ncu_summary.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/rNN_kernels.md"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram thr %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX thr %"), ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefronts % of peak"),
        ("lts__t_sectors.sum", "L2 sectors"), ("sm__cycles_elapsed.max", "cycles")]
SKIP = ("k_level_params", "k_grad_norm", "k_advance")
STALLS = ["long_scoreboard", "short_scoreboard", "mio_throttle", "lg_throttle", "barrier", "wait", "no_instructions", "math_pipe_throttle", "not_selected"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


print("| report | kernel | " + " | ".join(k for _, k in KEYS) + " | top stalls (pc samples) |")
print("|---|---|" + "---:|" * len(KEYS) + "---|")
for rep in sys.argv[1:]:
    hdr, units, data = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    for d in data:
        name = d[ix["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        if any(k in name for k in SKIP) or (ix.get("launch__grid_size") is not None and d[ix["launch__grid_size"]] in ("1", "2", "4")):
            continue
        cells = []
        for k, _ in KEYS:
            if k in ix and d[ix[k]] not in ("", "n/a"):
                v = d[ix[k]].replace(",", "")
                try:
                    f = float(v)
                    cells.append((f"{f:.3g}" if f < 1000 else f"{f:.0f}") + (" " + units[ix[k]] if units[ix[k]] not in ("%", "", "register/thread") else ""))
                except ValueError:
                    cells.append(v)
            else:
                cells.append("-")
        st = []
        for s in STALLS:
            k = f"smsp__pcsamp_warps_issue_stalled_{s}"
            if k in ix and d[ix[k]] not in ("", "n/a"):
                st.append((float(d[ix[k]].replace(",", "")), s))
        tot = sum(v for v, _ in st) or 1.0
        top = ", ".join(f"{s} {100 * v / tot:.0f}%" for v, s in sorted(st, reverse=True)[:4])
        print(f"| {rep.split('/')[-1]} | `{name}` | " + " | ".join(cells) + f" | {top} |")
