"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list (run here, no GPU):
    python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/rN_launch_summary_<what>.csv"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1], errors='replace') if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    ni = hdr.index('Metric Name')
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi or r[ni] != 'gpu__time_duration.sum':
            continue
        v = float(r[vi].replace(',', ''))
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
        name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', r[ki])
        name = re.sub(r'\(.*', '', name)[:110]
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    w = csv.writer(sys.stdout)
    w.writerow(['kernel', 'launches', 'device_time_us', 'share_pct'])
    w.writerow(['TOTAL', sum(cnt.values()), f'{total:.1f}', '100.0'])
    for k, v in tot.most_common(60):
        w.writerow([k, cnt[k], f'{v:.1f}', f'{100 * v / total:.2f}'])


if __name__ == '__main__':
    main()
