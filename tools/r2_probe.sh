mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --workload fixeddb --db-profiles 2500 --db-reads 1000 --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pfam shard', d['phases_ms_rank0'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['merged_hits'])"
