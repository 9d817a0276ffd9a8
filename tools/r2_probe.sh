mkdir -p gpurun_out/r2
for shape in 16,6,1 14,8,1 16,8,1; do
  DCPGPU_FORCE_SHAPE=$shape python bench.py --workload long --profiles 8 --reads 200 --steps 1 --warmup 1 --no-cpu --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$shape', d['phases_ms_rank0'], 'hits', d['hits_per_step_rank0'], 'value', round(d['value'],1))"
done
DCPGPU_FORCE_SHAPE=1,7,8 python bench.py --core 200 --profiles 64 --reads 1000 --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/s7_plain.log 2>&1 && DCPGPU_FORCE_SHAPE=1,7,8 ncu --set full --clock-control none --import-source on -k regex:k_score -s 1 -c 1 -o gpurun_out/r2/s7 python bench.py --core 200 --profiles 64 --reads 1000 --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/s7_ncu.log 2>&1
tail -1 gpurun_out/r2/s7_ncu.log
