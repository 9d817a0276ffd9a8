# traceback: parity tests, then phase times of a config-2-shaped, the config-2 and a config-4-shaped workload; with a
# variant build (-DDCP_TRACE_PROF=1) in gpurun_variants/lib_tprof.so also the cycles per phase of a hit
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary"
for w in "--profiles 50 --reads 4000" "" "--workload long --profiles 8 --reads 200"; do
  DCPGPU_TRACE_TIMES=1 $B $w > gpurun_out/r2/tr_probe.log 2> gpurun_out/r2/tr_probe.err
  grep "kernels done" gpurun_out/r2/tr_probe.err | tail -1
  tail -1 gpurun_out/r2/tr_probe.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['phases_ms_rank0'], d['merged_hits'])"
  if [ -f gpurun_variants/lib_tprof.so ]; then DCPGPU_LIB=$PWD/gpurun_variants/lib_tprof.so python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary $w 2>&1 | grep "prof\]" | tail -1; fi
done
