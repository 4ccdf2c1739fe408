#!/bin/bash
# round 2, session 20 (2 GPUs, short): torchrun bench at N=2 on a 1 M-passage world — harness check for the driver's SCALE run
# (file-system rendezvous of the cached world, NCCL init, calibration all_reduce, one all_gather per call)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s20
export LB2_BENCH_CHUNKS=1000000
t0=$(date +%s)
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --budget-s 30 > $O.n2.json 2> $O.n2.err; echo "n2 rc=$? wall=$(( $(date +%s) - t0 ))s"
grep "\[bench\]" $O.n2.err | tail -8
grep -m3 -i "nranks\|NVLS\|via P2P" $O.n2.err
python -c "
import json;d=json.load(open('$O.n2.json'));print({k:d[k] for k in ('value','recall_at_10','n_gpus','ms_per_step')}, d['e2e'], d['config']['parallelism'], d['clocks'])"
