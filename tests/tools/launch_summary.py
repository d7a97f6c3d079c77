"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals of the LAST step and,
with --nconv, every nconv_tc launch of it.  Usage: launch_summary.py file.csv [launches_per_step] [--nconv]"""
import csv, re, sys, collections

def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))

def short(n):
    n = re.sub(r'^void ', '', n)
    n = re.sub(r'\(.*', '', n)
    n = n.replace('gwn::', '').replace('tc::', '')
    return n[:70]

if __name__ == "__main__":
    rows = load(sys.argv[1])
    # a step starts at train_begin_kernel
    starts = [i for i, r in enumerate(rows) if 'train_begin_kernel' in r['Kernel Name']]
    last = rows[starts[-1]:] + []
    if len(starts) >= 2:
        last = rows[starts[-2]:starts[-1]]
    tot = collections.OrderedDict()
    for r in last:
        k = short(r['Kernel Name'])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r['Metric Value']) / 1e3
    s = sum(v[1] for v in tot.values())
    print(f"launches {len(last)}  serialised sum {s:.1f} us")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.1f} us {100 * v[1] / s:5.1f}%  x{v[0]:3d}  {k}")
    if '--nconv' in sys.argv:
        for r in last:
            if 'nconv_tc' in r['Kernel Name']:
                print(r['ID'], r['Grid Size'], float(r['Metric Value']) / 1e3)
