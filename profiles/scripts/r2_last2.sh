timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.1fM ms %.4f e2e %.1fM launches %d' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['gpu_launches']))"
