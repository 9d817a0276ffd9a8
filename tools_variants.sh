#!/bin/bash
# Experiment helper: time the score kernel for every kernel-variant library in gpurun_variants/.
for lib in gpurun_variants/lib_*.so; do
  for rep in 1 2; do
  DCPGPU_LIB=$PWD/$lib python bench.py --steps 2 --warmup 1 --no-cpu --profiles ${P:-200} --reads ${R:-4000} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', 'kernel_gcups=%.1f'%d['roofline']['kernel_gcups'], 'value=%.1f'%d['value'], d['phases_ms_rank0'])"
  done
done
