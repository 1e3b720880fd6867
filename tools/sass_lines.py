"""Static SASS instruction count per source line for one kernel of the in-tree .so (no GPU needed).
   python tools/sass_lines.py <cubin-substring> <mangled-substring> [line_lo line_hi]"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(root, 'boxinstseg_b200/lib/libboxseg_b200.so')], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if sys.argv[1] in f and f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and sys.argv[2] in l)
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 10**9)
cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
cur = ('', 0)
n = 0
for l in dis[start + 1:]:
    if l.startswith('.text.') or l.startswith('\t.section'):
        if n: break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m:
        n += 1
        cnt[cur] += 1
        ops[cur][m.group(2).split('.')[0]] += 1
src = {}
def text(f, ln):
    if f not in src:
        p = os.path.join(root, 'boxinstseg_b200/csrc', f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[f][ln - 1].strip()[:70] if 0 < ln <= len(src[f]) else ''
tot = 0
for (f, ln), c in sorted(cnt.items(), key=lambda kv: kv[0]):
    if 'onepass' in f and lo <= ln <= hi or (len(sys.argv) <= 4):
        pass
    if f.startswith('boxinst_onepass') and lo <= ln <= hi:
        tot += c
        print(f'{ln:4d} {c:4d}  {" ".join(f"{k}:{v}" for k, v in ops[(f, ln)].most_common(6)):50s} {text(f, ln)}')
print('total in range', tot, 'of', n)
