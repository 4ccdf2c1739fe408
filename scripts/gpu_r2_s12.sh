#!/bin/bash
# round 2, session 12: L2 prefetch distance of the GEMM producers (A/B), GEMM tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s12
python -c "from leann_b200 import build; build.needs_build() and build.build()" > $O.build.log 2>&1
for pf in 12 0 6 24 48; do
  LB2_GEMM_PF=$pf timeout 300 python scripts/kernel_bench.py > $O.kb_pf$pf.log 2>&1; echo "== pf=$pf rc=$?"; grep "gemm" $O.kb_pf$pf.log
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k gemm > $O.tests_a.log 2>&1; echo "tests_a rc=$?"; tail -2 $O.tests_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_ws_kernel -c 1 -o $O.gemm_ws python scripts/kernel_bench.py > $O.ncu_ws.log 2>&1; echo "ncu rc=$?"
