"""Developer tool: condenses one `ncu --set full --import-source on` capture of trace_kernel into the JSON kept under
profiles/ (key metrics, per-SASS-region instruction shares, stall samples).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_x_ncu_summary.json --label "..." --command "..."
"""
import argparse, csv, io, json, subprocess

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("out")
    ap.add_argument("--label", default="trace_kernel"); ap.add_argument("--command", default=""); ap.add_argument("--workload", default="")
    a = ap.parse_args()

    rows = list(csv.reader(io.StringIO(ncu_csv(a.rep, "raw"))))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    metrics = {}
    for n, u, v in zip(names, units, vals):
        if n in METRICS:
            metrics[n] = {"unit": u, "value": v}

    src = list(csv.reader(io.StringIO(ncu_csv(a.rep, "source"))))
    h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    col = {n: i for i, n in enumerate(src[h])}
    ins = []
    for r in src[h + 1:]:
        if len(r) <= col["Instructions Executed"]:
            continue
        try:
            ex = float(r[col["Instructions Executed"]] or 0); th = float(r[col["Thread Instructions Executed"]] or 0)
            smp = float(r[col["# Samples"]] or 0)
        except ValueError:
            continue
        ins.append((r[col["Source"]].split()[0] if r[col["Source"]] else "?", ex, th, smp))
    total = sum(e for _, e, _, _ in ins) or 1.0
    regions, start = [], 0
    for i in range(1, len(ins) + 1):
        if i == len(ins) or abs(ins[i][1] - ins[start][1]) > 0.05 * max(ins[start][1], 1.0):
            ex = sum(e for _, e, _, _ in ins[start:i]); th = sum(t for _, _, t, _ in ins[start:i])
            if ex / total >= 0.01:
                regions.append(f"sass[{start}-{i - 1}] n={i - start} exec/inst={ins[start][1]:.3e} warp%={100 * ex / total:.1f} "
                               f"lanes={th / max(ex, 1):.1f} samples={sum(s for _, _, _, s in ins[start:i]):.0f} first={ins[start][0]}")
            start = i
    mix = {}
    for op, ex, _, _ in ins:
        key = op.split(".")[0]
        mix[key] = mix.get(key, 0.0) + ex
    top_mix = {k: round(100 * v / total, 2) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:16]}

    stalls = {}
    for n, v in zip(names, vals):
        if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
            try:
                stalls[n.replace("smsp__pcsamp_warps_issue_stalled_", "stall_")] = int(float(v))
            except ValueError:
                pass
    stalls = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10])

    json.dump({"kernel": a.label, "command": a.command, "workload": a.workload, "metrics": metrics, "source_page_regions": regions,
               "warp_instruction_mix_pct": top_mix, "stall_samples": stalls}, open(a.out, "w"), indent=1)
    t = metrics.get("gpu__time_duration.sum", {})
    print("wrote", a.out, t)


if __name__ == "__main__":
    main()
