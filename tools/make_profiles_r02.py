"""tools/make_profiles_r02.py <capture-tag> [out-tag] — condenses the ncu captures of tools/gpu_profile_c3.sh (gpurun_out/<tag>_c3_{w1raw,w1,w8}.ncu-rep
and the launch lists <tag>_l_*.csv) into the tracked files under profiles/: <out>_c3_kernels.md, <out>_launches_c3_*.csv, and the
"c3" section of roofline_traffic.json (DRAM bytes per launch and kernel, what bench.py reports as `traffic`).  Reads reports with
`ncu -i`; needs no GPU.  The library on disk must be the build that was profiled (source-line tables)."""
import csv, io, json, os, shutil, sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_profiles as mp   # kernel_table / stall_table / lines_table

ROOT, OUT, PROF = mp.ROOT, mp.OUT, mp.PROF
SHAPES = (("w1raw", "whole frame on one GPU, raw-stream front (`clusters` = 0; what an unpartitioned submission takes by default)", ""),
          ("w1", "whole frame on one GPU, cluster front (`clusters` = 1)", ""),
          ("w8", "one rank's share of 8 (interleaved tile rows), cluster front", "@8"))


def dram_by_kernel(rep):
    hdr, units, data = mp.raw_table(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").strip()
        rd = float(r[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
        wr = float(r[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
        out[name] = int(rd + wr)
    return out


def main():
    tag = sys.argv[1]
    out_tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
    doc = [f"# {out_tag}: general-path kernels on C3 (1 M icosahedrons, 20 M triangles, 3840x2160)", "",
           "Captured by `tools/gpu_profile_c3.sh` (`ncu --set full --clock-control none`; per-launch times are cold-cache and",
           "serialised — use them for shares and counters; throughput numbers are the CUDA-event timings of `bench.py`).", ""]
    traffic = {}
    for shape, title, suffix in SHAPES:
        rep = os.path.join(OUT, f"{tag}_c3_{shape}.ncu-rep")
        if not os.path.exists(rep):
            continue
        out, *_ = mp.kernel_table(rep, f"C3, {title}")
        doc += out
        kernels = ("triangle_classify", "shade_tiles") if shape == "w1raw" else ("cluster_cull", "cluster_front", "direct_walk", "shade_tiles")
        for k in kernels:
            doc += mp.stall_table(rep, k)
        if shape in ("w1raw", "w1"):
            for k in (("triangle_classify",) if shape == "w1raw" else ("cluster_front", "direct_walk", "shade_tiles")):
                doc += mp.lines_table(rep, k, 16)
        doc.append("")
        d = dram_by_kernel(rep)
        if shape == "w1raw":      # N = 1 takes the raw-stream front by default
            traffic.update(d); traffic["frame"] = sum(d.values())
        elif shape == "w8":
            traffic.update({k + suffix: v for k, v in d.items()}); traffic["frame" + suffix] = sum(d.values())
        else:
            traffic.update({k + "@cluster": v for k, v in d.items() if k not in traffic})
        src = os.path.join(OUT, f"{tag}_l_{shape}.csv")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(PROF, f"{out_tag}_launches_c3_{shape}.csv"))
    open(os.path.join(PROF, f"{out_tag}_c3_kernels.md"), "w").write("\n".join(doc) + "\n")
    path = os.path.join(PROF, "roofline_traffic.json")
    rec = json.load(open(path)) if os.path.exists(path) else {}
    if "tile_raster_dram_bytes_per_launch" in rec:   # round 1's flat record -> the "c2" section
        rec = {"c2": {"tile_raster": {"dram_bytes_per_launch": rec["tile_raster_dram_bytes_per_launch"], "poses_per_launch": rec.get("poses_per_launch", 8),
                                      "issue_active_frac": 0.68, "source": "profiles/r01_c2_tile_raster.md"}}}
    rec["c3"] = dict(traffic, source=f"profiles/{out_tag}_c3_kernels.md: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full); "
                                     "plain names: whole frame on one GPU; '@8': one rank's share of 8; '@cluster': whole frame through the cluster front")
    json.dump(rec, open(path, "w"), indent=1)
    print(open(os.path.join(PROF, f"{out_tag}_c3_kernels.md")).read()[:3000])


if __name__ == "__main__":
    main()
