# cycle breakdown of the traceback per hit (variant build with -DDCP_TRACE_PROF=1)
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary"
for w in "--profiles 50 --reads 4000" "--workload long --profiles 8 --reads 200" "--core 512 --profiles 20 --reads 2000"; do
  DCPGPU_LIB=$PWD/gpurun_variants/lib_tprof.so DCPGPU_TRACE_TIMES=1 $B $w 2>&1 | grep "prof\]\|kernels done" | tail -2
done
