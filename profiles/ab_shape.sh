#!/bin/bash
cd "$(dirname "$0")/.."
MACM_SHAPE=16 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout.py tests/test_golden.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -5
for S in 2 1; do
MACM_SHAPE=16 python bench.py --no-legs --steps 500 --streams $S --e2e-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('16x4 streams',d['config']['streams'],'us/step %.2f'%(1e3*d['ms_per_step']), d['config']['launch'])"
done
