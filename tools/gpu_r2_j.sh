#!/bin/bash
# round 2, GPU call J: SM-cycle timeline of the tap kernel's roles (normal and all-knocked-out builds)
mkdir -p gpurun_out; O=gpurun_out
STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 120 python tools/tap_cycles_probe.py > $O/j_cycles.txt 2>&1
echo "=========== knocked out (no loads / MMAs / epilogue math / stores)" >> $O/j_cycles.txt
STGCN_B200_LIB=$PWD/build/variants/tl_ko.so timeout 120 python tools/tap_cycles_probe.py >> $O/j_cycles.txt 2>&1
cat $O/j_cycles.txt
