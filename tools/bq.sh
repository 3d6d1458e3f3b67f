python bench.py --no-cpu --no-llr8 --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
