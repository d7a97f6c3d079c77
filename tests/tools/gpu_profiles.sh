set -x
exp() {  # export a capture to text (details + selected raw metrics) and drop the big report: gpurun_out/ is capped at 64 MiB
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
python tests/tools/tf32_peak.py > gpurun_out/tf32_peak.log 2>&1; cat gpurun_out/tf32_peak.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_fp32x3.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-roofline --skip-tiers --no-graph > gpurun_out/ncu_l.log 2>&1
B="python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline --skip-tiers --no-graph"
ncu --set full --clock-control none --import-source on -k regex:nconv_tc_kernel -s 31 -c 2 -o gpurun_out/full_nconv -f $B > gpurun_out/ncu_f1.log 2>&1; exp full_nconv
ncu --set full --clock-control none --import-source on -k regex:tcpos_kernel -s 45 -c 3 -o gpurun_out/full_tcpos -f $B > gpurun_out/ncu_f2.log 2>&1; exp full_tcpos
ncu --set full --clock-control none --import-source on -k regex:tcred_kernel -s 16 -c 3 -o gpurun_out/full_tcred -f $B > gpurun_out/ncu_f3.log 2>&1; exp full_tcred
ncu --set full --clock-control none --import-source on -k regex:nconv_tc_kernel -s 2 -c 1 -o gpurun_out/full_nconv_2048 -f python tests/tools/configs_bench.py contraction_only > gpurun_out/ncu_f4.log 2>&1; exp full_nconv_2048
ls -la gpurun_out | head -40
