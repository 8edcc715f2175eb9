"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA load / store), UTCBAR (tcgen05.commit).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt      (runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'aivc_b200', 'libaivc_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
pat = ('UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'HMMA', 'FFMA')
counts, cur, order, k = {}, None, [], 0
for line in sass.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = re.sub(r'\(anonymous namespace\)::', '', names[k]).split('(')[0]
        k += 1
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur:
        for p in pat:
            if re.search(r'\b' + p + r'\b|\b' + p + r'\.', line):
                counts[cur][p] += 1
print('libaivc_b200.so (sm_100a), SASS mnemonic counts per kernel; built from this tree with nvcc 12.9')
print('%-58s' % 'kernel' + ''.join('%9s' % p for p in pat))
tot = collections.Counter()
for n in order:
    c = counts[n]
    tot.update(c)
    if any(c[p] for p in pat):
        print('%-58s' % n[:57] + ''.join('%9d' % c[p] for p in pat))
print('%-58s' % 'TOTAL' + ''.join('%9d' % tot[p] for p in pat))
