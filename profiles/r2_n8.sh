#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c_topo8.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n8.json 2> gpurun_out/r2c_bench_n8.err; echo bench rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 profiles/pcie_concurrent.py > gpurun_out/r2c_pcie8.json 2>/dev/null; echo pcie rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 profiles/pcie_concurrent.py --numa > gpurun_out/r2c_pcie8_numa.json 2>/dev/null; echo pcie-numa rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 profiles/peer_gather_check.py 2>&1 | tail -1
cat gpurun_out/r2c_pcie8.json gpurun_out/r2c_pcie8_numa.json
