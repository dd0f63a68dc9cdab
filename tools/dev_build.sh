#!/bin/sh
# Development build of libnmpc_b200.so with a single warp-path instantiation (default Nr = 6) and extra -D flags:
#   tools/dev_build.sh out.so [-DSOLVE_WARPS=4 -DSOLVE_MIN_CTAS=4 ...]
# Used for launch-configuration experiments (copy the variant over lib/libnmpc_b200.so on the GPU box); never shipped.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG="$ROOT/nonlinear-mpc-for-collision-free-and-deadlock-free-navigation-of-multiple-nonholonomic-mobile-robots_b200"
OUT=$1; shift
exec nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -I"$ROOT/include" -I"$PKG/csrc" \
     -DNMPC_DEV_ONLY_NR=${NMPC_DEV_NR:-6} "$@" -o "$OUT" "$PKG/csrc/nmpc_b200.cu"
