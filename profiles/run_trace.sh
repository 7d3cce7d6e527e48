#!/bin/bash
O=gpurun_out; mkdir -p $O
KCRN_GATHER=1,2 KCRN_LIBRARY=$PWD/profiles/microbench/_build/libkcrn_trace.so timeout 200 python profiles/microbench/variant_time.py ${1:-rcv1} > $O/r2q_sweep_trace.txt 2>&1
cat $O/r2q_sweep_trace.txt | cut -c1-200
