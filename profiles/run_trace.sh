#!/bin/bash
# %globaltimer phase marks of the sweep kernels (-DKCRN_SWEEP_TRACE build): Lanczos passes (slots 6, 7) next to the plain passes (20, 21)
O=gpurun_out; mkdir -p $O
TAG=${2:-r2w}
KCRN_GATHER=${KG:-1,2} KCRN_LIBRARY=$PWD/profiles/microbench/_build/libkcrn_trace.so timeout 200 python profiles/microbench/variant_time.py ${1:-rcv1} > $O/${TAG}_sweep_trace.txt 2>&1
cut -c1-200 $O/${TAG}_sweep_trace.txt | grep -v "slot  [0-589]\|slot 1[0-9]"
