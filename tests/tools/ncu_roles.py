"""Which barrier does each warp role of a warp-specialised kernel wait on?  Reads the source page of an ncu report
(`ncu --set full --import-source on`, built with -lineinfo), sums the stall samples per SASS instruction and prints the
hot ones with the source line they (and the code just before them) belong to: the `@P BRA` of a try_wait spin loop
shows up with the line of the barrier lambda it was inlined from, so "producer waits on empty 75 %" or "epilogue
warps wait on efull 52 %" can be read off directly (samples per warp ~ total / warps per CTA).
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > src.csv
       python tests/tools/ncu_roles.py src.csv [min_fraction=0.008]"""
import csv, sys

path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
rows = list(csv.reader(open(path)))
print([r[1][:100] for r in rows if r and r[0] == "Function Name"][:1])
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
seen = {}
for k, h in enumerate(hdrs):
    f = rows[h - 2][1] if rows[h - 2][0] == "File Path" else "?"
    si = rows[h].index("# Samples")
    end = hdrs[k + 1] - 2 if k + 1 < len(hdrs) else len(rows)
    cur = None
    for r in rows[h + 1:end]:
        if len(r) <= si:
            continue
        if r[0]:
            cur = (f.split("/")[-1], int(r[0]))
        a = r[2]
        if a and a != "-":
            try:
                av = int(a, 16) if a.startswith("0x") else int(a)
                n = int(r[si])
            except ValueError:
                continue
            seen.setdefault(av, [r[3][:80], 0, cur])
            seen[av][1] = max(seen[av][1], n)     # the same instruction is listed once per source file section
base = min(seen)
tot = sum(v[1] for v in seen.values())
print("total samples", tot)
prev = None
for a in sorted(seen):
    s, n, cur = seen[a]
    if cur and not cur[0].startswith(("tc_common", "common")):
        prev = cur
    if n >= tot * thr:
        print(f"{a - base:6x} {n:6d} {100 * n / tot:5.1f}%  {s:60s} | {cur[0]}:{cur[1]}  (kernel code before it: {prev})")
