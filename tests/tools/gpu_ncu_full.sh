# full ncu capture of the reduction kernels and of the node contraction (one step, no graph, no PDL)
export GWNET_B200_PDL=0 GWNET_B200_SIDE_STREAM=0
CMD="python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-roofline --skip-tiers --no-graph"
ncu --set full --clock-control none --import-source on -k regex:tcred_kernel -s 6 -c 6 -o gpurun_out/prof_tcred $CMD > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nconv_tc_kernel -s 30 -c 4 -o gpurun_out/prof_nconv $CMD > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/*.ncu-rep
