"""Per-source-line instruction / stall-sample totals of one kernel from an .ncu-rep (SASS page) joined with
nvdisasm line info of the in-tree .so.   python tools/ncu_lines.py <rep> <kernel-regex> <cubin-substring>"""
import csv, collections, os, re, subprocess, sys, tempfile
rep, kre, cub = sys.argv[1], sys.argv[2], sys.argv[3]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(root, 'boxinstseg_b200/lib/libboxseg_b200.so')], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if cub in f and f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kname = rows[0][1]
mangled_hint = re.search(r'::(\w+)<', kname).group(1)
targs = re.findall(r'\((int|bool)\)(\d+)', kname.split('(const')[0])
hdr = rows[1]
body = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    body.append(dict(zip(hdr, r)))
# locate the function in the disassembly
want = mangled_hint + 'I' + ''.join(('Li' if k == 'int' else 'Lb') + f'{t}E' for k, t in targs) if targs else mangled_hint
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and want in l)
lines, cur = [], ('', 0)
for l in dis[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        if lines: break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
        lines.append(cur)
print(f'{kname[:80]}\n sass rows ncu={len(body)} nvdisasm={len(lines)}')
n = min(len(body), len(lines))
inst, samp = collections.Counter(), collections.Counter()
for i in range(n):
    inst[lines[i]] += int(body[i]['Instructions Executed'] or 0)
    samp[lines[i]] += int(body[i]['# Samples'] or 0)
srcs = {}
def text_of(key):
    f, ln = key
    if f not in srcs:
        path = os.path.join(root, 'boxinstseg_b200/csrc', f)
        srcs[f] = [l.rstrip() for l in open(path)] if os.path.exists(path) else []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ''
tot, ts = sum(inst.values()), sum(samp.values())
print(f' total warp-instr {tot}, samples {ts}')
order = samp if (len(sys.argv) > 5 and sys.argv[5] == 'stall') else inst
for ln, _c in order.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 30):
    c = inst[ln]
    print(f'{c:9d} {100*c/tot:5.1f}%  stall {100*samp[ln]/max(ts,1):5.1f}%  {ln[0][:16]:16s}:{ln[1]:4d}: {text_of(ln)}')
