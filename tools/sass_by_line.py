"""Static SASS instruction count of one kernel, attributed to source file (and optionally line ranges).
usage: sass_by_line.py dis.txt kernel_substr [file_substr]   (dis.txt = nvdisasm --print-line-info cubin)"""
import re, sys, collections
txt = open(sys.argv[1]).read().split('\n'); kern = sys.argv[2]; fsel = sys.argv[3] if len(sys.argv) > 3 else None
infn = False; cur = ('?', 0); byfile = collections.Counter(); byline = collections.Counter()
for ln in txt:
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m: infn = kern in m.group(1); continue
    if ln.lstrip().startswith('.section'): infn = False; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', ln): byfile[cur[0]] += 1; byline[cur] += 1
print(sum(byfile.values()), 'instructions'); 
for f, c in byfile.most_common(): print(f'{c:7d} {f}')
if fsel:
    b = collections.Counter()
    for (f, l), c in byline.items():
        if fsel in f: b[l // 20 * 20] += c
    for l in sorted(b): print(f'  {fsel} lines {l:4d}-{l+19:4d}: {b[l]}')
