#!/bin/bash
# A/B of the convoy (NMPC_CONVOY = 0 off, 1 iteration barrier, 2 + barrier before the forward pass) for CTA shapes 3x4 (default), 4x3, 2x6;
# build the variants first: nvcc ... -DSOLVE_WARPS=4 -DSOLVE_MIN_CTAS=3 -o variants/w4.so (and 2 / 6 -> variants/w2.so)
PKG=nonlinear-mpc-for-collision-free-and-deadlock-free-navigation-of-multiple-nonholonomic-mobile-robots_b200
cp lib/libnmpc_b200.so /tmp/default.so
for v in default w4 w2; do
  if [ $v = default ]; then cp /tmp/default.so lib/libnmpc_b200.so; else cp variants/$v.so lib/libnmpc_b200.so; fi
  for c in 0 2 2; do
    echo "== $v convoy=$c"; NMPC_CONVOY=$c timeout 300 python tools/time_batch.py 6 8192 | tail -1
  done
done
cp /tmp/default.so lib/libnmpc_b200.so
