for v in 0 2; do echo "== BHMC_SCHEDULE=$v"; BHMC_SCHEDULE=$v python bench.py --steps 8 --warmup 3 --no-ess --no-cpu-baseline --no-pixels --blocks "" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f e2e %.0f launches %s' % (d['value'], d['e2e']['value'], d['gpu_launches']))"; done
