# ncu --set full of the traceback kernel on a small config-2-shaped workload (4000 hits, 50 profiles)
mkdir -p gpurun_out/r2
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --profiles 50 --reads 4000"
DCPGPU_TRACE_TIMES=1 $B 2>&1 | grep "kernels\|phases" | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -f -o gpurun_out/r2/tr_c2b $B > gpurun_out/r2/tr_c2b_ncu.log 2>&1
