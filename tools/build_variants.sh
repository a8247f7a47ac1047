#!/bin/bash
# Compile-time A/B builds of the library: tools/build_variants.sh NAME "-DFLAG=V ..." -> build/variants/libgraphtap_b200.NAME.so
# (select with GT_LIB=<path>; build/ is git-ignored but travels with gpurun)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; FLAGS=$2
W=$(mktemp -d)
mkdir -p $W/graphtap_b200/csrc $W/include $ROOT/build/variants
cp $ROOT/graphtap_b200/csrc/*.cu $ROOT/graphtap_b200/csrc/*.cuh $ROOT/graphtap_b200/csrc/*.h $ROOT/graphtap_b200/csrc/*.cpp $ROOT/graphtap_b200/csrc/Makefile $W/graphtap_b200/csrc/
cp -r $ROOT/include/* $W/include/
make -C $W/graphtap_b200/csrc -j4 EXTRA="$FLAGS" OUT=$ROOT/build/variants/libgraphtap_b200.$NAME.so > /dev/null
rm -rf $W
ls -la $ROOT/build/variants/libgraphtap_b200.$NAME.so
