#!/bin/bash
mkdir -p gpurun_out
for lib in libspimcuda.so libspimcuda_minb7.so libspimcuda_minb8.so; do
  echo "== $lib"; SPIMCUDA_LIB=$PWD/spimagine_b200/$lib EXP_ISO_VARIANTS=4:1 timeout 300 python scripts/exp_iso.py 2>&1 | grep -v "min/max\|hit rays"
done 2>&1 | tee gpurun_out/exp_iso_minb.txt
for c in 3 8 12 16; do
  echo "== occ ctas/SM $c"; EXP_OCC_CTAS=$c EXP_ISO_VARIANTS=4:1 timeout 300 python scripts/exp_iso.py 2>&1 | grep "full chain"
done 2>&1 | tee gpurun_out/exp_occ_ctas.txt
