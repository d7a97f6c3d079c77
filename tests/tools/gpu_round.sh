set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tf32.log 2>&1; tail -c 600 gpurun_out/bench_tf32.log
python bench.py --steps 10 --precision fp32x3 --skip-cpu-baseline > gpurun_out/bench_fp32x3.log 2>&1
python bench.py --steps 10 --precision fp32 --skip-cpu-baseline > gpurun_out/bench_fp32.log 2>&1
python tests/tools/ref_eager_gpu.py > gpurun_out/ref_eager_gpu.log 2>&1; cat gpurun_out/ref_eager_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_l.log 2>&1
for k in tcred_kernel nconv_tc_kernel RowMlp RowGateBwd RowSeg; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 2 -o gpurun_out/full_$k -f python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_full_$k.log 2>&1
done
ls -la gpurun_out
