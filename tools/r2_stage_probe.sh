# 4/5-nt emission lines through shared memory (cp.async) against the LDG path, at 8 and 7 nodes per lane, M = 200
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --profiles 200 --reads 4000"
run() { echo "== $1"; env $1 $B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('kernel_gcups=%.1f value=%.1f' % (d['roofline']['kernel_gcups'], d['value']), d['phases_ms_rank0'], d['merged_hits'])"; }
run "X=0"
run "DCPGPU_TMA=2"
run "DCPGPU_FORCE_SHAPE=1,7,8"
run "DCPGPU_FORCE_SHAPE=1,7,8 DCPGPU_TMA=2"
