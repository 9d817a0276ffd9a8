#!/bin/bash
# Experiment helper: time the score kernel for every kernel-variant library in gpurun_variants/.
# usage: P=profiles R=reads CORES="200 256" ./tools_variants.sh
for core in ${CORES:-200}; do
for lib in gpurun_variants/lib_*.so; do
  DCPGPU_LIB=$PWD/$lib python bench.py --steps 2 --warmup 1 --no-cpu --profiles ${P:-200} --reads ${R:-4000} --core $core 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('core=$core', '$lib', 'kernel_gcups=%.1f'%d['roofline']['kernel_gcups'], 'value=%.1f'%d['value'], d['phases_ms_rank0'])"
done
done
