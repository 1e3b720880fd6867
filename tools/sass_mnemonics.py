"""Static evidence (no GPU): tensor-core / TMA / async-copy / cluster mnemonics per kernel of the in-tree library.
   python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'boxinstseg_b200', 'lib', 'libboxseg_b200.so')
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
cur, counts, samples = None, collections.OrderedDict(), {}
pat = re.compile(r'\b(UTCHMMA[\w.]*|UTMALDG[\w.]*|LDTM[\w.]*|UTCBAR[\w.]*|SYNCS[\w.]*|UBLKCP[\w.]*|LDGSTS[\w.]*|UTCATOMSWS[\w.]*|UTMAPF[\w.]*|UCGABAR[\w.]*)')
for line in out.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        samples[cur] = {}
        continue
    if cur:
        m = pat.search(line)
        if m:
            op = m.group(1)
            counts[cur][op] += 1
            samples[cur].setdefault(op, re.sub(r'/\*[^*]*\*/', '', line).strip())
print('# cuobjdump -sass boxinstseg_b200/lib/libboxseg_b200.so : tensor-core / TMA / async-copy / cluster mnemonics per kernel')
print('# (count, mnemonic, first occurrence).  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld,')
print('# UBLKCP = cp.async.bulk (1-D TMA), LDGSTS = cp.async, SYNCS = mbarrier ops, UCGABAR = cluster barrier')
for k, c in counts.items():
    if not c:
        continue
    print('\n' + re.sub(r'^_ZN3bxs\d+_GLOBAL__N__[0-9a-f_]+_cu_[0-9a-f]+', '', k))
    for op, n in sorted(c.items(), key=lambda x: -x[1]):
        print(f'  {n:4d}  {op:32s} {samples[k][op][:110]}')
