for cfg in "256 128" "128 128" "128 64" "256 64"; do set -- $cfg; SCFLOW_FC_KR0=$1 SCFLOW_FC_KR1=$2 python bench.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('kr0/kr1 $1/$2', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['breakdown']['per_iter_ms'],4))"; done
