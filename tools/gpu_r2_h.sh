#!/bin/bash
# round 2, GPU call H: knock-out probe of the tap kernel (which resource bounds the tile period) + one ncu --set full capture
mkdir -p gpurun_out; O=gpurun_out
KO_TAG=base timeout 120 python tools/ko_probe.py > $O/h_ko.txt 2>&1
for v in ko_store ko_epi ko_load ko_mma ko_se ko_sel ko_all; do
  KO_TAG=$v STGCN_B200_LIB=$PWD/build/variants/$v.so timeout 120 python tools/ko_probe.py >> $O/h_ko.txt 2>&1
done
cat $O/h_ko.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:umma_tap_kernel -c 6 -f -o $O/h_tap python tools/ko_probe.py > $O/h_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O
