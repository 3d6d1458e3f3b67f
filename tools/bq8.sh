# quick end-to-end lines (16-bit, 8-bit LLRs): tools/bq8.sh <label> [bench args]
lab=$1; shift
python bench.py --no-cpu "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lab', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e8', round(d['e2e_llr8']['value']), round(d['e2e_llr8']['ms_per_step'],3), 'tb', round(d['e2e_tb']['value']))"
