# 256-bit emission loads ([lane][8] line layout) against the product layout, by kernel class (core length -> class)
V=$PWD/gpurun_variants/lib_e256.so
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --profiles 100 --reads 3000"
for core in 192 384 448 96 768 576 320 80 1024 2048; do for lib in "" $V; do
  DCPGPU_LIB=$lib $B --core $core 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('core=$core lib=${lib##*/} kernel_gcups=%.1f' % (d['roofline']['kernel_gcups']), d['merged_hits']['digest'])"
done; done
