"""Splits the SASS page of an .ncu-rep (ncu --set full --import-source on) into the regions between BAR.SYNC instructions and
prints each region's share of the warp-stall samples and of the executed instructions with its dominant opcodes.

    python tools/ncu_regions.py gpurun_out/x.ncu-rep
"""
import csv
import io
import subprocess
import sys


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[isamp]) for r in data)
    totex = sum(int(r[iex]) for r in data)
    print(f'{rows[0][1]}: {len(data)} SASS instructions, {tot} samples, {totex} warp instructions executed')
    regions, cur = [], []
    for r in data:
        cur.append(r)
        if 'BAR.SYNC' in r[isrc]:
            regions.append(cur)
            cur = []
    regions.append(cur)
    for i, g in enumerate(regions):
        s = sum(int(r[isamp]) for r in g)
        e = sum(int(r[iex]) for r in g)
        if s < 0.003 * tot:
            continue
        ops = {}
        for r in g:
            t = r[isrc].split()
            op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            ops[op] = ops.get(op, 0) + int(r[iex])
        top = ' '.join(f'{k}:{v / max(e, 1):.0%}' for k, v in sorted(ops.items(), key=lambda x: -x[1])[:5])
        st = {}
        for c in stall_cols:
            st[hdr[c]] = sum(int(r[c] or 0) for r in g)
        stt = ' '.join(f'{k[6:]}:{v / max(s, 1):.0%}' for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f'region {i:2d}: {len(g):4d} instr  samples {100 * s / tot:5.1f}%  executed {100 * e / totex:5.1f}%  | {top} | {stt}')


if __name__ == '__main__':
    main()
