"""tools/make_profiles.py <round-tag> — turns the ncu captures a GPU session left in gpurun_out/ into the tracked
summaries under profiles/ (reads reports with `ncu -i`; needs no GPU).

  gpurun_out/<tag>_c2_tile_raster.ncu-rep        -> profiles/<tag>_c2_tile_raster.md  (+ roofline_traffic.json)
  gpurun_out/<tag>_c3_w1.ncu-rep / _w8.ncu-rep   -> profiles/<tag>_c3_kernels.md       (classify, shade, vertex stage)
  gpurun_out/<tag>_c4.ncu-rep                    -> profiles/<tag>_c4_kernels.md
  gpurun_out/l_w1.csv, l_w8.csv, l_c4.csv        -> profiles/<tag>_launches_*.csv      (per-launch time lists)
"""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency (threads/inst)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
]


def raw_table(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2:]


def kernel_table(rep, title):
    hdr, units, data = raw_table(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    names = [r[ix["Kernel Name"]].split("(")[0] for r in data]
    out = [f"### {title}", "", f"`{os.path.basename(rep)}` (`ncu --set full --clock-control none`; cold-cache, serialised replays)", "",
           "| metric | unit | " + " | ".join(names) + " |", "|---|---|" + "---|" * len(data)]
    for key, label in METRICS:
        if key in ix:
            out.append(f"| {label} | {units[ix[key]]} | " + " | ".join(r[ix[key]] for r in data) + " |")
    return out, hdr, units, data


def stall_table(rep, kernel):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    try:
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    except StopIteration:
        return []
    hdr = rows[hi]
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    tot = {hdr[i]: 0 for i in cols}
    for r in rows[hi + 1:]:
        if r and r[0] == "Address":
            break
        for i in cols:
            if i < len(r) and r[i].isdigit():
                tot[hdr[i]] += int(r[i])
    s = sum(tot.values()) or 1
    out = ["", f"Warp-state samples, `{kernel}`:", "", "| state | share |", "|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        out.append(f"| {k} | {100 * v / s:.1f} % |")
    return out


def lines_table(rep, kernel, top=18):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kernel, str(top)],
                         capture_output=True, text=True).stdout
    return ["", f"Warp instructions by source line, `{kernel}` (tools/ncu_lines.py):", "", "```", txt.rstrip(), "```"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    rep = os.path.join(OUT, f"{tag}_c2_tile_raster.ncu-rep")
    if os.path.exists(rep):
        out, hdr, units, data = kernel_table(rep, "C2 (bench workload): `tile_raster`, 3840x2160, data.bin scene")
        seg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_segments.py"), rep, "tile_raster"],
                             capture_output=True, text=True).stdout
        out = [f"# {tag}: tile_raster on the bench workload (one launch = 8 poses: the capture passes `--views-per-launch 8`)", ""] + out + stall_table(rep, "tile_raster") + \
            ["", "Warp time barrier to barrier (tools/ncu_segments.py):", "", "```", seg.rstrip(), "```"] + lines_table(rep, "tile_raster", 30)
        open(os.path.join(PROF, f"{tag}_c2_tile_raster.md"), "w").write("\n".join(out) + "\n")
        ix = {h: i for i, h in enumerate(hdr)}
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(data[0][ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
        wr = float(data[0][ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
        grid = float(data[0][ix["launch__grid_size"]]) if "launch__grid_size" in ix else 4080.0
        json.dump({"tile_raster_dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                   "poses_per_launch": int(round(grid / 4080.0)),   # 60 x 68 tiles per 3840x2160 pose
                   "source": f"profiles/{tag}_c2_tile_raster.md (ncu --set full, one launch of `poses_per_launch` poses: the DRAM writes are "
                             "the frames themselves, 33 MB each, minus what is still in the 126 MB L2 when the launch ends)"},
                  open(os.path.join(PROF, "roofline_traffic.json"), "w"), indent=1)
    doc = [f"# {tag}: general-path kernels on C3 (1 M icosahedrons, 20 M triangles, 4K)", ""]
    for w, title in (("w1", "whole frame on one GPU"), ("w8", "one rank's share of 8 (interleaved tile rows)")):
        rep = os.path.join(OUT, f"{tag}_c3_{w}.ncu-rep")
        if os.path.exists(rep):
            out, *_ = kernel_table(rep, f"C3, {title}")
            doc += out
            for k in ("triangle_classify", "shade_tiles"):
                doc += stall_table(rep, k)
            if w == "w1":
                for k in ("triangle_classify", "shade_tiles"):
                    doc += lines_table(rep, k)
            doc.append("")
    if len(doc) > 2:
        open(os.path.join(PROF, f"{tag}_c3_kernels.md"), "w").write("\n".join(doc) + "\n")
    rep = os.path.join(OUT, f"{tag}_c4.ncu-rep")
    if os.path.exists(rep):
        out, *_ = kernel_table(rep, "C4 (clipping stress, 1 M triangles, 34 % straddle), whole frame on one GPU")
        doc = [f"# {tag}: general-path kernels on C4", ""] + out
        for k in ("post_setup", "tile_raster_queue"):
            doc += stall_table(rep, k) + lines_table(rep, k, 12)
        open(os.path.join(PROF, f"{tag}_c4_kernels.md"), "w").write("\n".join(doc) + "\n")
    for src, dst in (("l_w1.csv", f"{tag}_launches_c3_whole_frame.csv"), ("l_w8.csv", f"{tag}_launches_c3_one_of_8.csv"),
                     ("l_c4.csv", f"{tag}_launches_c4.csv"), (f"launches_c2_{tag}.csv", f"{tag}_launches_bench_c2_48frames.csv")):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
    par = os.path.join(OUT, f"parity_{tag}")
    if os.path.isdir(par):
        shutil.rmtree(os.path.join(PROF, f"parity_{tag}"), ignore_errors=True)
        shutil.copytree(par, os.path.join(PROF, f"parity_{tag}"))
    if os.path.exists(os.path.join(OUT, "configs.jsonl")):
        latest = {}   # the most recent single-GPU line of every configuration (C1..C5)
        for l in open(os.path.join(OUT, "configs.jsonl")):
            if l.strip() and '"gpus": 1' in l:
                latest[json.loads(l)["config"][:2]] = l
        if latest:
            open(os.path.join(PROF, f"{tag}_configs_1gpu.jsonl"), "w").writelines(latest[k] for k in sorted(latest))
    for src in (f"bench_{tag}.json", f"bench_ref_{tag}.json"):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, src))


if __name__ == "__main__":
    main()
