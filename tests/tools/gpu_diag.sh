run() {
  echo "== $1"
  env $1 timeout 300 python bench.py --steps 50 --warmup 3 --skip-cpu-baseline --skip-tiers 2>&1 | tail -1 > gpurun_out/diag.json
  python - <<'P'
import json
d = json.loads(open('gpurun_out/diag.json').read())
print(d['ms_per_step'])
print('  '.join(f"{o['op']}={o['ms_per_step']*1e3:.0f}" for o in d.get('operators', [])))
P
}
run "GWNET_B200_L2PROMO=128"
run "GWNET_B200_L2PROMO=256"
run "GWNET_B200_L2PROMO=0"
run "GWNET_B200_L2PROMO=128"
