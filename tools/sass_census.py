"""Per-kernel opcode census of the built library: which kernels use the 5th-generation tensor cores (UTC*MMA), tensor memory
(LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), the legacy warp-level MMA path (HMMA) and cp.async (LDGSTS).

    python tools/sass_census.py > profiles/sass_census.txt        # needs only cuobjdump (no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, '2g-gcn_b200', 'lib2ggcn_b200.so')
PATTERNS = [('UTC*MMA', r'\bUTC[A-Z]*MMA'), ('UTCBAR', r'\bUTCBAR'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'), ('UTMALDG', r'\bUTMALDG'),
            ('UTMASTG', r'\bUTMASTG'), ('UBLKCP', r'\bUBLKCP'), ('SYNCS', r'\bSYNCS'), ('HMMA', r'\bHMMA'), ('LDGSTS', r'\bLDGSTS'),
            ('FFMA', r'\bFFMA')]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    print(f'# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): occurrences of selected opcodes per kernel')
    print(f'# {"kernel":70s} ' + ' '.join(f'{n:>8s}' for n, _ in PATTERNS))
    for fn, c in counts.items():
        name = re.sub(r'\(.*', '', demangle(fn).replace('(anonymous namespace)::', '').replace('tg::', ''))
        print(f'{name[:72]:72s} ' + ' '.join(f'{c.get(n, 0):8d}' for n, _ in PATTERNS))


if __name__ == '__main__':
    main()
