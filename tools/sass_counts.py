"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libspi_b200.so (runs without a GPU):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync.      python tools/sass_counts.py > profiles/rN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'spi_b200', 'libspi_b200.so')
PAT = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMAPF', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'HMMA', 'RED.E', 'ATOMG']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'\(anonymous namespace\)::', '', name)
            cur = re.sub(r'\(.*', '', name)[:90]
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in PAT:
            if re.search(r'\b' + re.escape(p), line):
                counts[cur][p] += 1
    print('kernel | ' + ' | '.join(PAT))
    tot = collections.Counter()
    for k, c in counts.items():
        if any(c[p] for p in PAT[:11]) or c['HMMA']:
            print(k + ' | ' + ' | '.join(str(c[p]) for p in PAT))
        tot.update(c)
    print('TOTAL (all %d kernels) | ' % len(counts) + ' | '.join(str(tot[p]) for p in PAT))


if __name__ == '__main__':
    main()
