set -x
python -m pytest tests/test_gpu_reset.py -q 2>&1 | tail -5
for S in 1 2 3 4; do python bench.py --no-legs --steps 500 --streams $S --e2e-steps 12 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('streams',d['config']['streams'],'ms',d['ms_per_step'])"; done
MACM_BLOCK_THREADS=448 python bench.py --no-legs --steps 500 --streams 2 --e2e-steps 12 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('448x2 streams',d['config']['streams'],'ms',d['ms_per_step'])"
for S in 1 2; do python bench.py --no-legs --envs 32768 --rot 2 --steps 100 --streams $S --e2e-steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5 streams',d['config']['streams'],'ms',d['ms_per_step'], d['config']['contacts_per_agent'])"; done
python bench.py --no-legs --envs 32768 --rot 4 --steps 100 --streams 1 --settle 200 --e2e-steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5 rot4 settle200 streams',d['config']['streams'],'ms',d['ms_per_step'], d['config']['contacts_per_agent'])"
