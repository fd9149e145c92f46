"""tools/ncu_segments.py <report.ncu-rep> <kernel-substring> — where a kernel's warps spend their time, barrier to barrier.

Splits the SASS of an `ncu --set full` capture at every BAR.SYNC and prints, per segment, its share of the warp
stall samples and of the executed warp instructions, the barrier-stall samples charged to its first instructions
(= warps waiting for the slowest warp of the PREVIOUS segment) and the source line range the segment maps to.
A segment with few instructions but many samples (its own or as barrier wait after it) runs on too few warps.
Development aid; reads reports, needs no GPU."""
import csv, io, os, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import line_table  # noqa: E402


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {k: i for i, k in enumerate(hdr)}
    data = []
    for r in rows[h + 1:]:
        if r and r[0] == "Address":
            break
        if len(r) == len(hdr):
            data.append(r)
    first = int(data[0][0], 16)
    table = line_table(kernel)
    tot_s = sum(int(r[ix["# Samples"]]) for r in data) or 1
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
    print(f"{kernel}: {tot_s} samples, {tot_i} warp instructions")
    print(f"{'segment':>17} {'samples %':>10} {'inst %':>8} {'barrier wait %':>15}  kernels.cu lines")
    seg = {"s": 0, "i": 0, "b": 0, "start": 0, "lines": set()}

    def flush(end):
        ls = sorted(l for f, l in seg["lines"] if f == "kernels.cu")
        span = f"{ls[0]}-{ls[-1]}" if ls else "-"
        print(f"{seg['start']:#8x}-{end:#8x} {100 * seg['s'] / tot_s:>10.1f} {100 * seg['i'] / tot_i:>8.1f} {100 * seg['b'] / tot_s:>15.1f}  {span}")

    for r in data:
        off = int(r[0], 16) - first
        if "BAR.SYNC" in r[ix["Source"]]:
            flush(off)
            seg = {"s": 0, "i": 0, "b": 0, "start": off, "lines": set()}
        seg["s"] += int(r[ix["# Samples"]]); seg["i"] += int(r[ix["Instructions Executed"]]); seg["b"] += int(r[ix["stall_barrier"]])
        if int(r[ix["Instructions Executed"]]) and off in table:
            seg["lines"].add(table[off])
    flush(int(data[-1][0], 16) - first)


if __name__ == "__main__":
    main()
