#!/bin/bash
# block width x streams on the bench workload (500-step windows): does retiring envs in smaller groups beat one
# lock-step 896-thread block per SM once two or four batches are in flight?
cd "$(dirname "$0")/.."
for T in 896 448 224 128; do for S in 2 4; do
  MACM_BLOCK_THREADS=$T python bench.py --no-legs --steps 500 --streams $S --e2e-steps 0 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads $T streams $S: %.2f us/step, blocks/SM %d' % (1e3*d['ms_per_step'], d['config']['launch']['blocks_per_sm']))"
done; done
