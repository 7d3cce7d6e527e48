#!/bin/bash
O=gpurun_out; mkdir -p $O
for f in 0 4 6 7; do
echo "=== KCRN_PERSIST_FLAGS=$f"
KCRN_PERSIST_FLAGS=$f KCRN_PERSIST_TRACE=1 timeout 300 python profiles/microbench/persist_probe.py rcv1 2>&1 | grep -E "kcrn:  |persistent  " | head -23
done > $O/r2e_flags.txt 2>&1
cat $O/r2e_flags.txt
