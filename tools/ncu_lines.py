"""tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [top] [stall] — warp instructions per CUDA source line (sorted by
instructions, or by stall samples with a fourth argument `stall`).

Joins the SASS page of an `ncu --set full` capture (instructions executed, thread instructions, stall samples per
SASS instruction) with the line table of the shipped cubin (`nvdisasm --print-line-info`; the library is built
with -lineinfo), by instruction offset inside the kernel.  The library on disk must be the build that was
profiled.  Development aid; reads reports, needs no GPU."""
import csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def line_table(kernel: str):
    tmp = tempfile.mkdtemp(prefix="cubin_")
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "swift3drenderer_b200", "lib", "render.so")],
                          cwd=tmp, stdout=subprocess.DEVNULL)
    table = {}
    for name in os.listdir(tmp):
        if not name.startswith("kernels.") or not name.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, name)], capture_output=True, text=True).stdout
        inside, cur = False, None
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                inside = kernel in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
            if m and cur:
                table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[hdr_i + 1:]:
        if r and r[0] == "Address":
            break  # a second launch of the same kernel follows: keep the first
        if len(r) > ix["Thread Instructions Executed"] and r[ix["Instructions Executed"]].isdigit():
            data.append(r)
    base = int(data[0][0], 16)
    table = line_table(kernel)
    src = {}
    agg = {}
    total = 0
    for r in data:
        off = int(r[0], 16) - base
        key = table.get(off, ("?", 0))
        ie, te, smp = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])
        a = agg.setdefault(key, [0, 0, 0, 0])
        a[0] += ie; a[1] += te; a[2] += smp; a[3] += 1
        total += ie
    tot_smp = sum(a[2] for a in agg.values()) or 1
    print(f"{kernel}: {total} warp instructions, {len(data)} SASS instructions, {len(agg)} source lines")
    print(f"{'file:line':<22}{'warp inst':>12}{'%':>7}{'lanes':>7}{'stall %':>9}{'sass':>6}  source")
    by = 2 if len(sys.argv) > 4 and sys.argv[4] == "stall" else 0
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
        fn, ln = key
        if fn not in src:
            path = os.path.join(ROOT, "swift3drenderer_b200", "csrc", fn)
            src[fn] = open(path).read().splitlines() if os.path.exists(path) else []
        text = src[fn][ln - 1].strip()[:90] if 0 < ln <= len(src[fn]) else ""
        print(f"{fn + ':' + str(ln):<22}{a[0]:>12}{100 * a[0] / total:>7.1f}{a[1] / max(a[0], 1):>7.1f}{100 * a[2] / tot_smp:>9.1f}{a[3]:>6}  {text}")


if __name__ == "__main__":
    main()
