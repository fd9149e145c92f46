"""tools/ncu_summary.py <report.ncu-rep> [out.md] — condenses an `ncu --set full` capture into the few
numbers DESIGN.md / bench.py cite (reads the report with `ncu -i`, needs no GPU)."""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs, CTAs/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem, CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency (threads/inst)"),
    ("smsp__thread_inst_executed_per_inst_executed.pct", "warp execution efficiency %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu summary of `{rep.split('/')[-1]}`", "",
           f"{len(data)} launch(es) of `{data[0][ix['Kernel Name']]}`; values per launch.", "",
           "| metric | unit | " + " | ".join(f"launch {k}" for k in range(len(data))) + " |",
           "|---|---|" + "---|" * len(data)]
    for key, label in KEYS:
        if key in ix:
            out.append(f"| {label} (`{key}`) | {units[ix[key]]} | " + " | ".join(r[ix[key]] for r in data) + " |")
    # stall reasons from the source page (SASS), summed over the kernel
    try:
        sass = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], text=True,
                                       stderr=subprocess.DEVNULL)
        srows = list(csv.reader(io.StringIO(sass)))
        h = srows[1]
        six = {c: i for i, c in enumerate(h)}
        body = [r for r in srows[2:] if len(r) >= len(h)]
        stalls = {c: 0.0 for c in h if c.startswith("stall_") and "Not Issued" not in c}
        for r in body:
            for c in stalls:
                try:
                    stalls[c] += float(r[six[c]])
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        out += ["", "Warp-state samples over the first profiled launch (SASS source page):", ""]
        out += ["| state | share |", "|---|---|"]
        for c, v in sorted(stalls.items(), key=lambda kv: -kv[1]):
            if v / tot >= 0.005:
                out.append(f"| {c} | {100 * v / tot:.1f} % |")
    except subprocess.CalledProcessError:
        pass
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
