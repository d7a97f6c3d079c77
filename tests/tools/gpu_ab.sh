# A/B of the launch-level switches: PDL on/off x nconv column split forced 1 / auto.  One JSON line each.
B="python bench.py --steps 100 --warmup 3 --skip-cpu-baseline --skip-roofline"
for pdl in 0 1; do for nwt in 1 0; do
  echo "== PDL=$pdl NWT=$nwt"
  GWNET_B200_PDL=$pdl GWNET_B200_NCONV_NWT=$nwt timeout 300 $B 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print(d['ms_per_step'], d['value'], d.get('other_tiers', {}).get('tf32', {}).get('ms_per_step'))
except Exception as e:
    print('FAILED', l[-600:])
"
done; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
