mkdir -p gpurun_out/r2
DCPGPU_FORCE_SHAPE=1,7,8 python tools/sanitize_probe.py 193 200 210 224 2>&1 | tail -1
DCPGPU_PIPE=1 python tools/sanitize_probe.py 129 150 161 190 193 256 2>&1 | tail -1
for cfg in "1,7,8 200 224" ; do set -- $cfg; DCPGPU_FORCE_SHAPE=$1 python tools/class_sweep.py $2 $3 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('p7', d['M'], d['padded'], d['score_ms'], d['gcups'], d['padded_gnodes_per_s'])"; done
DCPGPU_PIPE=1 python tools/class_sweep.py 160 192 200 256 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('pipe', d['M'], d['q'], d['padded'], d['score_ms'], d['gcups'], d['padded_gnodes_per_s'])"
