"""Summarise `ncu --page raw --csv` exports (tools/prof_round.sh) into a markdown table of the counters the roofline discussion uses,
one column per captured launch, and (with --traffic out.json) the per-launch DRAM bytes keyed by bench timer name.
usage: ncu_csv_summary.py [--traffic profiles/ncu_traffic.json] raw.csv [raw.csv ...] > summary.md"""
import csv
import json
import re
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active cycles)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("smsp__inst_executed.sum", "warp instructions"),
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
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def short(name):
    n = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    return n.split("(")[0]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return {h: i for i, h in enumerate(hdr)}, units, data


def table(path):
    ix, units, data = load(path)
    names = [short(r[ix["Kernel Name"]]) for r in data]
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


def dram_bytes(ix, units, r):
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1.0)
    return tot


def traffic(paths):
    """bench timer name -> DRAM bytes of ONE launch (mean over the captured launches that map to the timer)."""
    acc = {}
    for path in paths:
        ix, units, data = load(path)
        seen = {}
        for r in data:
            n = short(r[ix["Kernel Name"]])
            k = seen[n] = seen.get(n, 0) + 1
            t = None
            m = re.match(r"conv_tma3?_kernel<(\d+), (\d+)>", n)
            if m:                                     # first launch of a shape = forward, second = data gradient (of the layer that has the swapped shape)
                ci, co = int(m.group(1)), int(m.group(2))
                layer = {(20, 20): 2, (20, 40): 3, (40, 40): 4, (40, 20): 3}[(ci, co)]
                t = f"conv{layer}_fwd" if (k == 1 and (ci, co) != (40, 20)) else f"conv{layer}_dgrad"
            m = re.match(r"conv_wgrad_tma_kernel<(\d+), (\d+)>", n)
            if m:
                t = f"conv{ {(20, 20): 2, (20, 40): 3, (40, 40): 4}[(int(m.group(1)), int(m.group(2)))] }_wgrad"
            m = re.match(r"planes_kernel<(\d+), (\d)>", n)
            if m:
                c, mode = int(m.group(1)), int(m.group(2))
                if mode == 0:
                    t = "conv2_planes" if (c == 20 and k == 1) else "conv3_planes" if c == 20 else "conv4_planes"
                else:
                    t = "conv2_dy_planes" if c == 20 else ("conv4_dy_planes" if k == 1 else "conv3_dy_planes")
            if n.startswith("conv1_fwd"):
                t = "conv1_fwd"
            if n.startswith("conv1_wgrad"):
                t = "conv1_wgrad"
            if n.startswith("decm_fwd_kernel"):
                t = "note_decoder_fwd"
            if n.startswith("decm_bwd_kernel"):
                t = "note_decoder_bwd"
            if n.startswith("gru_seq_fwd"):
                t = "encoder_gru_fwd"
            if n.startswith("gru_seq_bwd"):
                t = "encoder_gru_bwd"
            if t:
                acc.setdefault(t, []).append(dram_bytes(ix, units, r))
    return {k: sum(v) / len(v) for k, v in acc.items()}


if __name__ == "__main__":
    args = sys.argv[1:]
    out = None
    if args and args[0] == "--traffic":
        out, args = args[1], args[2:]
    for p in args:
        table(p)
    if out:
        json.dump(traffic(args), open(out, "w"), indent=1, sort_keys=True)
