#!/bin/bash
# Round 2, call AW: why the e2e leg of the full default bench (9.6-9.9 M) is below the short bench and the tool (13 M): same box, same call.
set -x
python bench.py --no-cpu-baseline --no-other-configs --policy random 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full steps (2000)', d['e2e']['by_mode'])"
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-other-configs --policy random 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('200 steps', d['e2e']['by_mode'])"
python bench.py --steps 2000 --warmup 50 --no-cpu-baseline --no-other-configs --policy random --cuda-graph off 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('2000 steps no graph', d['e2e']['by_mode'])"
python tools/e2e_ab.py --threads 16 --steps 20 2>&1 | cut -c1-150
python bench.py --no-cpu-baseline --no-other-configs --policy random --e2e-steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full steps, e2e 200 steps', d['e2e']['by_mode'])"
