#!/bin/bash
# bench.py (cold / warm / e2e / latency) with and without the convoy, default CTA shape and 6x2
PKG=nonlinear-mpc-for-collision-free-and-deadlock-free-navigation-of-multiple-nonholonomic-mobile-robots_b200
cp $PKG/libnmpc_b200.so /tmp/default.so
for v in default w6; do
  if [ $v = default ]; then cp /tmp/default.so $PKG/libnmpc_b200.so; else cp variants/$v.so $PKG/libnmpc_b200.so; fi
  for c in 0 1; do
    echo "== $v convoy=$c"; NMPC_CONVOY=$c timeout 300 python bench.py --swarm 0 | tail -1
  done
done
cp /tmp/default.so $PKG/libnmpc_b200.so
