# ncu --set full of the three tcgen05 kernel families; the reports are turned into text on the box (they exceed the
# 64 MiB that gpurun copies back) and removed
export GWNET_B200_PDL=0 GWNET_B200_SIDE_STREAM=0
CMD="python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline --skip-tiers --no-graph"
KEYS='Kernel Name|Grid Size|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|lts__throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|lts__t_sector_hit_rate.pct'
cap() {   # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/ncu_full_$1.details.txt 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > /tmp/raw_$1.csv 2>/dev/null
  python - $1 "$KEYS" <<'P'
import csv, json, sys
name, keys = sys.argv[1], sys.argv[2].split('|')
rows = list(csv.reader(open('/tmp/raw_%s.csv' % name)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
out = [{k + (' [' + units[idx[k]] + ']' if units[idx[k]] else ''): d[idx[k]] for k in keys if k in idx} for d in data]
json.dump(out, open('gpurun_out/ncu_key_%s.json' % name, 'w'), indent=1)
P
  rm -f /tmp/prof_$1.ncu-rep
}
cap nconv nconv_tc_kernel 30 4
cap tcred tcred_kernel 6 6
cap tcpos tcpos_kernel 40 8
ls -la gpurun_out/
