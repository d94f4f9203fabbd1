#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 120 python tools/fb2_cycles_probe.py > $O/p_fb2_cycles.txt 2>&1
cat $O/p_fb2_cycles.txt
