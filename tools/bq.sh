# quick device-resident bench line: tools/bq.sh <label> [bench args]
lab=$1; shift
python bench.py --no-cpu --no-llr8 --no-e2e "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lab', round(d['value']), round(d['ms_per_step'],4))"
