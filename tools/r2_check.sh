# GPU tests, then the default bench line (with the secondary shapes) at a few steps
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2/bench_check.json 2> gpurun_out/r2/bench_check.err; tail -c 300 gpurun_out/r2/bench_check.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_check.json').read().strip().splitlines()[-1])
print(d['value'], d['phases_ms_rank0'], d['e2e']['value'], d['merged_hits'])
for k,v in d['secondary'].items(): print(k, round(v['gcups'],1), round(v['score_ms'],1), round(v['trace_ms'],1), round(v['total_ms'],1), v['hits'], v['hits_digest'], v['oracle_sample'])
PY
