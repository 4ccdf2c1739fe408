#!/bin/bash
# round 2, session 6 (2 GPUs): torchrun bench at N=2 + the reference arm, 1 M passages (fast) — harness check for the driver's SCALE run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s6
python -c "from leann_b200 import build; build.needs_build() and build.build()" > $O.build.log 2>&1
export LB2_BENCH_CHUNKS=1000000
timeout 900 python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 --ref-budget-s 45 > $O.ref.json 2> $O.ref.err; echo "ref rc=$?"; tail -c 600 $O.ref.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --budget-s 60 > $O.n2.json 2> $O.n2.err; echo "n2 rc=$?"
grep "\[bench\]" $O.n2.err | tail -8; python -c "
import json;d=json.load(open('$O.n2.json'));print({k:d[k] for k in ('value','recall_at_10','n_gpus')}, d['e2e'], d.get('parity'), d['config']['parallelism'])"
timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 --budget-s 60 --no-cpu-baseline > $O.n1.json 2> $O.n1.err; echo "n1 rc=$?"
python -c "
import json;d=json.load(open('$O.n1.json'));print({k:d[k] for k in ('value','recall_at_10','n_gpus')}, d['e2e']['value'], d.get('parity'), d.get('cpu_baseline'))"
