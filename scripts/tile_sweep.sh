#!/bin/bash
# (stages, M-tiles per CTA) sweep of the TMA weight-gradient and implicit-GEMM kernels at the C2 batch:
#   bash scripts/tile_sweep.sh > gpurun_out/tile_sweep.txt
run() { echo -n "$1=$2  "; env "$1=$2" python scripts/time_layer.py $3 $4 $5 2>&1 | tail -n 1; }
echo "== wgrad, encoder conv1 (Cb=32, Cs=64, 262144 pixels)";  for c in 4,1 2,2 3,2 2,4 3,4; do run BN_WG64 $c 0 1 2; done
echo "== wgrad, encoder conv2 (Cb=64, Cs=128, 65536 pixels)";  for c in 3,1 2,2 3,2 2,4; do run BN_WG128 $c 0 2 2; done
echo "== wgrad, encoder conv3 (Cb=128, Cs=256, 16384 pixels)"; for c in 4,1 3,2; do run BN_WG256 $c 0 3 2; done
echo "== igemm forward, encoder conv1 (Co=64)";  for c in 4,1 2,2 3,2 4,2; do run BN_IG64 $c 0 1 0; done
echo "== igemm forward, encoder conv2 (Co=128)"; for c in 3,1 2,2 3,2 4,2; do run BN_IG128 $c 0 2 0; done
