"""Static loop bodies of one kernel: for every backward branch, the SASS instructions between its target label and the
branch, with their opcode mix and the source lines they come from (no GPU needed).
   python tools/sass_loops.py <cubin-substring> <mangled-substring> [min_instructions]"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get('BXS_SO', os.path.join(root, 'boxinstseg_b200/lib/libboxseg_b200.so'))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if sys.argv[1] in f and f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and sys.argv[2] in l)
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 20
items, labels, cur = [], {}, 0
for l in dis[start + 1:]:
    if l.startswith('.text.') or l.startswith('\t.section'):
        if items: break
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r'(\.L_x_\d+):', l)
    if m:
        labels[m.group(1)] = len(items); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);', l)
    if m:
        items.append((m.group(2), m.group(3), cur))
for i, (op, rest, ln) in enumerate(items):
    if op.startswith('BRA'):
        m = re.search(r'`\((\.L_x_\d+)\)', rest)
        if m and m.group(1) in labels and labels[m.group(1)] <= i:
            a = labels[m.group(1)]
            body = items[a:i + 1]
            if len(body) < minlen: continue
            lines = collections.Counter(x[2] for x in body)
            ops = collections.Counter(x[0].split('.')[0] for x in body)
            lo, hi = min(lines), max(lines)
            print(f'loop of {len(body)} instructions, source lines {lo}-{hi}')
            print('   ops:', ' '.join(f'{k}:{v}' for k, v in ops.most_common(14)))
            print('   by line:', ' '.join(f'{k}:{v}' for k, v in sorted(lines.items())))
