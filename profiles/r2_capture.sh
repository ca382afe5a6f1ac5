#!/bin/bash
# round-2 evidence run: GPU tests, the driver's bench command, then (each only after its command ran clean without
# ncu) the launch list of the bench and one full capture of a steady single-step launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err || { tail -5 gpurun_out/r2c_bench.err; exit 1; }
python bench.py --no-legs --steps 20 --warmup 5 --e2e-steps 0 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1024 -c 560 --csv --log-file gpurun_out/r2c_ncu_launches.csv \
    python bench.py --no-legs --steps 20 --warmup 5 --e2e-steps 0 > gpurun_out/r2c_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:macm_step_kernel --launch-skip 1100 -c 1 -f -o gpurun_out/r2c_steady \
    python bench.py --no-legs --steps 20 --warmup 5 --e2e-steps 0 >> gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out | tail -5
