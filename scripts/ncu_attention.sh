#!/bin/bash
# ncu --set full of the attention kernel on the micro-benchmark shapes (1 GPU)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 2 -o gpurun_out/attn \
    python scripts/kernel_bench.py 65536 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attention rc=$?"
