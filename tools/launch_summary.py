"""Condenses an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file x.csv) into the per-kernel table kept under profiles/.

    python tools/launch_summary.py gpurun_out/x.csv "what was run" [first_kernel_regex] > profiles/rNN_launches_summary.txt

With a third argument only the launches from the LAST occurrence of a kernel matching it onwards are counted (e.g. 'geo_bn_stats' =
the last training step of the capture).
"""
import collections
import csv
import re
import sys


def main():
    path, what = sys.argv[1], sys.argv[2]
    first = sys.argv[3] if len(sys.argv) > 3 else None
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    if first:
        idx = [i for i, r in enumerate(rows) if re.search(first, r['Kernel Name'])]
        rows = rows[idx[-1]:]
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows:
        name = re.sub(r'\(.*', '', r['Kernel Name'])[:70]
        key = (name, r.get('Grid Size', ''))
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        us = v / 1000 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000)
        tot[key] += us
        cnt[key] += 1
    total = sum(tot.values())
    print(f'# ncu launch list of {what} (ncu --metrics gpu__time_duration.sum --clock-control none)')
    print(f'# (cold-cache, serialised launches: compare SHARES, not absolutes).  {len(rows)} launches, total {total / 1000:.3f} ms')
    print(f'{"kernel":72s} {"grid":16s} {"count":>6s} {"mean_us":>10s} {"share%":>7s}')
    for key, us in tot.most_common():
        print(f'{key[0]:72s} {key[1]:16s} {cnt[key]:6d} {us / cnt[key]:10.1f} {100 * us / total:7.2f}')


if __name__ == '__main__':
    main()
