for cfg in "512 384" "384 288" "512 192"; do set -- $cfg; echo "=== threads $1 TR $2"; CORA_B200_PTHREADS=$1 CORA_B200_TILE_ROWS=$2 timeout 300 python bench.py --no-cpu-baseline --no-solve 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f  us/CG %.1f  e2e %.1f  frac %.3f grid %d' % (d['value'],d['us_per_cg_iteration'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['grid']))
print('  '.join('%s %.1f' % (k,v['avg_us']) for k,v in d['roofline']['phases_in_kernel_globaltimer_cta0'].items()))
"; done
