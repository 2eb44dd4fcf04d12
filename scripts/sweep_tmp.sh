TWISTERL_B200_PRECISION=f16x2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "balanced" 2>&1 | tail -5
run() { timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' value %.3e ms %.2f e2e %.3e'%(d['value'],d['ms_per_step'],d['e2e']['value']))"; }
echo bal0; TWISTERL_B200_BALANCE=0 run
echo bal2; run
echo bal4; TWISTERL_B200_BALANCE=4 run
echo bal1; TWISTERL_B200_BALANCE=1 run
