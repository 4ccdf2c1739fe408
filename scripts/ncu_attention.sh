#!/bin/bash
# ncu --set full of the attention kernel at L=128 on the micro-benchmark (1 GPU).
# kernel_bench launches attention 6x at L=64 first (1 warm-up + 5 timed): skip those.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 7 -c 1 -o gpurun_out/attn \
    python scripts/kernel_bench.py 262144 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attention rc=$?"
