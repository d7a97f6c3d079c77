# parity, then bench with / without the deferred weight gradients
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for dw in 0 1; do
echo "== DEFER_WGRAD=$dw"
GWNET_B200_DEFER_WGRAD=$dw timeout 300 python bench.py --steps 100 --warmup 3 --skip-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_quick_$dw.json
python - $dw <<'P'
import json, sys
d = json.loads(open('gpurun_out/bench_quick_%s.json' % sys.argv[1]).read())
print(d['ms_per_step'], d['value'], d.get('other_tiers', {}).get('tf32', {}).get('ms_per_step'), d.get('gpu_launches_per_step'))
for o in d.get('operators', []):
    print(f"{o['op']:22s} {o['ms_per_step']*1e3:8.1f} us  hbm {o['hbm_frac']:.2f}")
P
done
