#!/bin/bash
# round 2, session 8: attention A/B (empty-warp skip, sleeping waits), GEMM wait A/B, then the driver's own commands at 10 M:
# reference arm first, then this repo's arm, each timed by wall clock (limit 870 s per arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s8
python -c "from leann_b200 import build; build.build(force=True)" > $O.build.log 2>&1
timeout 300 python scripts/attn_debug.py > $O.attn_debug.log 2>&1; echo "attn_debug rc=$? nonfinite-lines: $(grep -c 'non-finite [1-9]' $O.attn_debug.log) bad-rows: $(grep -c 'rows>tol' $O.attn_debug.log)"
for v in "skip1:LB2_ATTN_SKIP=1" "skip0:LB2_ATTN_SKIP=0" "skip1w64:LB2_ATTN_WAIT_NS=64" "skip1w256:LB2_ATTN_WAIT_NS=256" "legacy:LB2_ATTN_LEGACY=1"; do
  n=${v%%:*}; e=${v#*:}
  env $e timeout 300 python scripts/kernel_bench.py > $O.kb_$n.log 2>&1; echo "== $n rc=$?"; grep "attention" $O.kb_$n.log
done
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -x -q > $O.tests_a.log 2>&1; echo "tests_a rc=$?"; tail -2 $O.tests_a.log
grep -i "gemm\|layer" $O.kb_skip1.log | head -30
LB2_GEMM_WAIT_NS=64 timeout 300 python scripts/kernel_bench.py > $O.kb_w64.log 2>&1; grep -i "gemm" $O.kb_w64.log | head -12
t0=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O.ref.json 2> $O.ref.err; echo "ref rc=$? wall=$(( $(date +%s) - t0 ))s"
tail -c 900 $O.ref.json; grep "\[bench\]" $O.ref.err | tail -12
t0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O.bench.json 2> $O.bench.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
grep "\[bench\]" $O.bench.err | tail -25
python - <<PY
import json
d=json.load(open('$O.bench.json'))
print({k:d.get(k) for k in ('value','recall_at_10','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d.get('parity')); print(d.get('cpu_baseline')); print(d.get('clocks')); print(d.get('breakdown'))
PY
