#!/bin/bash
# Run under gpurun on ONE GPU.  Produces, in gpurun_out/:
#   launches.csv   every kernel launch of a short bench run with its device time (ncu, serialised, cold cache)
#   gemm.ncu-rep   ncu --set full of 3 launches of the tcgen05 GEMM (one per epilogue type where possible)
# Set-up kernels (corpus embedding, torch graph builder) are skipped by name filter + launch-skip.
set -u
mkdir -p gpurun_out
CH=${CH:-60000}; Q=${Q:-256}
KREG='regex:gemm_f16_tn_kernel|attention_kernel|layernorm_kernel|embed_ln_kernel|attention_items_kernel|pool_kernel|hnsw_step_kernel|gather_bounds_kernel|init_slots_kernel'
# set-up launches of our own kernels: ceil(CH/4096) passes + query encode, 45 kernels each (MiniLM: 3 + 6*7)
SKIP=$(( ( (CH + 4095) / 4096 + 2 ) * 45 ))
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s $SKIP -c 4000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --chunks $CH --queries $Q --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tn_kernel -s $(( SKIP / 45 * 24 + 200 )) -c 4 \
    -o gpurun_out/gemm python bench.py --chunks $CH --queries $Q --steps 1 --warmup 0 --no-cpu-baseline \
    > gpurun_out/prof_bench2.json 2> gpurun_out/prof_bench2.err
echo "gemm capture rc=$?"
ls -la gpurun_out/
