# parity + one bench line (fp32x3 and tf32 tiers), no CPU baseline
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 100 --warmup 3 --skip-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_quick.json
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_quick.json').read())
print(d['ms_per_step'], d['value'], d.get('other_tiers', {}).get('tf32', {}).get('ms_per_step'))
for o in d.get('operators', []):
    print(f"{o['op']:22s} {o['ms_per_step']*1e3:8.1f} us  hbm {o['hbm_frac']:.2f}")
P
