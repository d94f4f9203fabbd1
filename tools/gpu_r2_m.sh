#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 120 python tools/tap_cycles_probe.py > $O/m_cycles.txt 2>&1
cat $O/m_cycles.txt
