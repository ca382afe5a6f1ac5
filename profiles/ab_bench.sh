#!/bin/bash
# A/B of libmacm builds on the bench workload: us/step with two streams and with one (500-step windows)
cd "$(dirname "$0")/.."
for lib in "$@"; do
  for S in 2 1; do
    MACM_LIB=$PWD/profiles/_variants/libmacm_$lib.so python bench.py --no-legs --steps 500 --streams $S --e2e-steps 0 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib streams',d['config']['streams'],'us/step %.2f'%(1e3*d['ms_per_step']))"
  done
done
