"""Dynamic instruction counts / stall samples per CUDA source line of one kernel, from an ncu report captured with
--import-source on (SASS page) and the object's line table (nvdisasm -g on the cubin of the same build).
usage: python tools/hot_lines.py report.ncu-rep object.o kernel_regex mangled_name source.cu [top]"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, obj, kregex, mangled, srcfile = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = glob.glob(tmp + '/*.cubin')[0]
dis = subprocess.run(['nvdisasm', '-g', cub], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith('.text.' + mangled + ':'))
lines, cur = [], None
for l in dis[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kregex],
                     capture_output=True, text=True).stdout
hdr, inst, seen = None, [], 0
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == 'Kernel Name':
        seen += 1
        hdr = None
        continue
    if r and r[0] == 'Address':
        hdr = r
        continue
    if hdr and seen == 1 and len(r) > 6:
        inst.append((int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('# Samples')]), r[1].strip()))
n = min(len(lines), len(inst))
print(f'{len(lines)} SASS instructions in the object, {len(inst)} in the report')
agg, sm = collections.Counter(), collections.Counter()
for (f, s, t), ln in zip(inst[:n], lines[:n]):
    agg[ln] += f
    sm[ln] += s
tot, ts = sum(agg.values()), sum(sm.values())
src = open(srcfile).read().splitlines()
base = os.path.basename(srcfile)
print(f'total warp instructions {tot}, samples {ts}')
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    text = src[ln[1] - 1].strip()[:100] if ln and ln[0] == base else ''
    print('%5.1f%% inst %5.1f%% samp %-24s %s' % (100 * v / tot, 100 * sm[ln] / max(ts, 1), '%s:%d' % ln if ln else '?', text))
