#!/bin/bash
# scripts/build_variant.sh NAME "EXTRA NVCC FLAGS" — builds build/NAME/csrc/liboppositerenderer_b200.so from a copy of the
# sources with extra -D flags, for A/B timing on the GPU box: OPR_LIBRARY=build/NAME/csrc/liboppositerenderer_b200.so python bench.py
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
D=$ROOT/build/$NAME/csrc
mkdir -p "$D" "$ROOT/build/include"
cp "$ROOT"/include/*.h "$ROOT/build/include/"
cp "$ROOT"/oppositerenderer_b200/csrc/*.cu "$ROOT"/oppositerenderer_b200/csrc/*.cuh "$ROOT"/oppositerenderer_b200/csrc/*.h \
   "$ROOT"/oppositerenderer_b200/csrc/*.cpp "$ROOT"/oppositerenderer_b200/csrc/Makefile "$D/"
make -s -j8 -C "$D" liboppositerenderer_b200.so EXTRA="$*"
echo "$D/liboppositerenderer_b200.so"
