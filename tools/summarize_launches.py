"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: ms per step and share per kernel.
usage: python tools/summarize_launches.py launches.csv <model steps in the run> [title]"""
import collections
import csv
import re
import sys

path, steps = sys.argv[1], int(sys.argv[2])
title = sys.argv[3] if len(sys.argv) > 3 else path
rows = list(csv.reader(open(path, errors='replace')))
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[start]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg, cnt = collections.Counter(), collections.Counter()
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[ui], 1e-6)
    name = re.sub(r'\(.*$', '', r[ki]).strip()
    agg[name] += v * scale
    cnt[name] += 1
tot = sum(agg.values())
own = sum(v for k, v in agg.items() if 'hwgat::' in k)
blas = sum(v for k, v in agg.items() if 'nvjet' in k or 'cublas' in k.lower() or 'cutlass' in k)
print(f'# {title}\n')
print(f'total {tot:.1f} ms over {sum(cnt.values())} launches = {tot / steps:.1f} ms per step; hwgat:: kernels '
      f'{100 * own / tot:.1f}%, cuBLAS / CUTLASS {100 * blas / tot:.1f}%, other PyTorch kernels '
      f'{100 * (tot - own - blas) / tot:.1f}%.\n')
print('| ms/step | share | launches/step | kernel |\n|---:|---:|---:|---|')
for k, v in agg.most_common(45):
    print(f'| {v / steps:.2f} | {100 * v / tot:.1f}% | {cnt[k] / steps:.1f} | `{k}` |')
