"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: our kernels one by one, library kernels pooled.
usage: python profiles/scripts/launch_summary.py <csv> <first-launch-index-of-pop_fg_kernel occurrence> [<end occurrence>]"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
names = [(r[4], float(r[-1].replace(',', ''))) for r in rows]
idx = [i for i, (n, _) in enumerate(names) if 'pop_fg_kernel' in n]
a = idx[int(sys.argv[2])] - 1
b = idx[int(sys.argv[3])] - 1 if len(sys.argv) > 3 and int(sys.argv[3]) < len(idx) else len(names)
tot = lib = nlib = 0
for n, t in names[a:b]:
    ours = any(t in n for t in ('sl::', 'tc::', 'tcg::', 'bwd::'))
    tot += t
    if ours:
        print(f'{t / 1000:9.1f} us  {re.sub(r"[(].*", "", n)[:100]}')
    else:
        lib += t; nlib += 1
print(f'total {tot / 1000:.1f} us; of which {nlib} PyTorch/cuBLAS parameter-side kernels {lib / 1000:.1f} us')
