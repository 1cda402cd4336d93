for cfg in 1; do echo "=== CORA_B200_REG=$cfg"; CORA_B200_REG=$cfg timeout 300 python bench.py --no-cpu-baseline --no-solve 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f  us/CG %.1f  e2e %.1f  frac %.3f' % (d['value'],d['us_per_cg_iteration'],d['e2e']['value'],d['roofline']['frac']))
print('  '.join('%s %.1f' % (k,v['avg_us']) for k,v in d['roofline']['phases_in_kernel_globaltimer_cta0'].items()))
"; done
