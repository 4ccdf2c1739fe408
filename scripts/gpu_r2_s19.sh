#!/bin/bash
# round 2, session 19 (final single-GPU pass): full GPU test suite, smoke(), the driver's bench command at 10 M, kernel
# microbench, then the ncu profiles (launch list + captures) of scripts/profile_r02.sh
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s19
t0=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu > $O.tests.log 2>&1; echo "tests rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -4 $O.tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O.smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O.smoke.log
t0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O.bench.json 2> $O.bench.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d=json.load(open('$O.bench.json'))
print({k:d.get(k) for k in ('value','recall_at_10','ms_per_step','gpu_launches')}); print(d['e2e']['value']); print(d['roofline']); print(d.get('parity')); print(d.get('clocks')); print({k:d['detail'][k] for k in ('encoder_share','attention_share','encoder_algorithmic_tflops')})
PY
timeout 300 python scripts/kernel_bench.py > $O.kb.log 2>&1; cat $O.kb.log | grep -v "^$"
bash scripts/profile_r02.sh
