"""Hottest SASS instructions (warp-stall samples) of every kernel launch in an .ncu-rep (ncu --set full --import-source on).

    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    for b in range(len(starts) - 1):
        blk = rows[starts[b]:starts[b + 1]]
        hdr = blk[1]
        data = [r for r in blk[2:] if len(r) == len(hdr)]
        isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(int(r[isamp]) for r in data)
        print(f'=== launch {b}: {blk[0][1]}: {len(data)} SASS instructions, {tot} samples')
        order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top_n]
        for i in sorted(order):
            r = data[i]
            st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
            stt = ' '.join(f'{n}:{v}' for v, n in st if v)
            print(f'{i:5d} {100 * int(r[isamp]) / max(tot, 1):5.1f}%  ex={r[iex]:>7}  {r[isrc].strip()[:70]:70s} {stt}')


if __name__ == '__main__':
    main()
