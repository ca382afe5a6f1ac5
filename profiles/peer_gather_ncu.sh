#!/bin/bash
# NVLink bytes of ONE PeerGather step launch on a shard rank (its peer stores into the learner's memory), from ncu's
# link counters.  Two ranks; both run under ncu (-c 1 each), no collective inside the profiled window.
cd "$(dirname "$0")/.."
M=nvltx__bytes.sum
timeout 240 ncu --target-processes all --metrics $M,gpu__time_duration.sum --clock-control none -k regex:macm_step_kernel --launch-skip 200 -c 1 \
   --csv --log-file gpurun_out/r2b_peer_gather_ncu.csv \
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 profiles/peer_gather_nvlink.py > gpurun_out/r2b_peer_gather_ncu.log 2>&1
echo rc=$?
grep -E "nvl|duration" gpurun_out/r2b_peer_gather_ncu.csv | cut -d, -f5,10,13-15
