"""prints gpurun_out/l_w{1,8}.csv (written by tools/c3_quick.sh) as one line per kernel"""
import csv
for w in (1, 8):
    rows = [r for r in csv.reader(open(f'gpurun_out/l_w{w}.csv')) if len(r) > 10]
    h = rows[0]; ix = {n: i for i, n in enumerate(h)}
    d = {}
    for r in rows[1:]:
        d.setdefault((r[ix['ID']], r[ix['Kernel Name']][:22]), {})[r[ix['Metric Name']]] = r[ix['Metric Value']]
    print('world', w)
    tot = 0
    for (i, k), m in d.items():
        tot += float(m['gpu__time_duration.sum'])
        print(' ', k.ljust(22), f"{float(m['gpu__time_duration.sum'])/1e3:8.1f} us", f"inst {float(m['smsp__inst_executed.sum'])/1e6:7.1f} M",
              f"issue {float(m['smsp__issue_active.avg.pct_of_peak_sustained_active']):5.1f} %", f"lanes {float(m['smsp__thread_inst_executed_per_inst_executed.ratio']):5.1f}",
              f"dram rd {float(m['dram__bytes_read.sum'])/1e6:7.1f} MB")
    print('  total', f"{tot/1e3:.1f} us")
