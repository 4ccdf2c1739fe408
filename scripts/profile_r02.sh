#!/bin/bash
# Round-2 profiles (ONE GPU, under gpurun): launch list of a short recompute search + ncu --set full captures of the three
# kernels the verdict asks about (tcgen05 GEMMs: weight-stationary + CTA-pair LayerNorm variant, tcgen05 attention, hnsw_step_kernel in recompute mode).
# Uses a 200 k-passage world so that set-up stays short; numbers printed by these runs are never bench values.
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r02prof
export LB2_BENCH_CHUNKS=200000
python bench.py --steps 1 --warmup 1 --queries 512 --no-cpu-baseline > $O.warm.json 2> $O.warm.err   # builds + caches the world
KREG='regex:gemm_f16_tn_kernel|gemm_f16_ws_kernel|gemm_f16_ln_kernel|gemm_f16_ln_pair_kernel|attention_tc_kernel|attention_tc_items_kernel|embed_ln_kernel|pool_kernel|hnsw_step_kernel|gather_bounds_kernel|init_slots_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 6000 --csv --log-file $O.launches.csv \
    python bench.py --steps 1 --warmup 0 --queries 512 --no-cpu-baseline > $O.l.json 2> $O.l.err; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_ws_kernel -s 40 -c 2 -o $O.gemm \
    python bench.py --steps 1 --warmup 0 --queries 512 --no-cpu-baseline > /dev/null 2> $O.g.err; echo "gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_ln_pair_kernel -s 40 -c 2 -o $O.gemm_ln \
    python bench.py --steps 1 --warmup 0 --queries 512 --no-cpu-baseline > /dev/null 2> $O.gl.err; echo "gemm_ln rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 10 -c 2 -o $O.attn \
    python bench.py --steps 1 --warmup 0 --queries 512 --no-cpu-baseline > /dev/null 2> $O.a.err; echo "attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_step_kernel -s 10 -c 2 -o $O.step \
    python bench.py --steps 1 --warmup 0 --queries 512 --no-cpu-baseline > /dev/null 2> $O.s.err; echo "step rc=$?"
ls -la gpurun_out/ | grep r02prof
