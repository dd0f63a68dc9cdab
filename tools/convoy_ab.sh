#!/bin/bash
# A/B of the convoy (NMPC_CONVOY = 0 off, 1 iteration barrier, 2 + barrier before the forward pass) for CTA shapes 4x3 and 6x2
PKG=nonlinear-mpc-for-collision-free-and-deadlock-free-navigation-of-multiple-nonholonomic-mobile-robots_b200
cp $PKG/libnmpc_b200.so /tmp/default.so
for v in default w3 w2; do
  if [ $v = default ]; then cp /tmp/default.so $PKG/libnmpc_b200.so; else cp variants/$v.so $PKG/libnmpc_b200.so; fi
  for c in 0 2 2; do
    echo "== $v convoy=$c"; NMPC_CONVOY=$c timeout 300 python tools/time_batch.py 6 8192 | tail -1
  done
done
cp /tmp/default.so $PKG/libnmpc_b200.so
