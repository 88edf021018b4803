"""Summarise `ncu --set full` reports (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into a markdown table of
the counters the roofline discussion uses, one column per captured launch.  usage: ncu_summary.py rep [rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active cycles)"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/CTA"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def table(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = []
    for r in data:
        n = r[ix["Kernel Name"]].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        names.append(n.split("(")[0])
    print(f"### `{path.split('/')[-1]}`\n")
    print("| counter | unit | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|---|" + "---:|" * len(names))
    for key, label in WANT:
        if key not in ix:
            continue
        i = ix[key]
        vals = []
        for r in data:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.4g}" if abs(v) < 1e6 else f"{v:.4e}")
            except ValueError:
                vals.append(r[i])
        print(f"| {label} (`{key}`) | {units[i]} | " + " | ".join(vals) + " |")
    print()
    return [(n, {k: (data[j][ix[k]], units[ix[k]]) for k, _ in WANT if k in ix}) for j, n in enumerate(names)]


if __name__ == "__main__":
    for p in sys.argv[1:]:
        table(p)
