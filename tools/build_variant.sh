#!/bin/bash
# Build a variant of libmsda_b200.so with extra -D flags for ONE source file (A/B experiments):
#   tools/build_variant.sh <name> <source.cu> <flags...>  ->  <pkg>/variants/libmsda_b200_<name>.so
# Select it at run time with MSDA_B200_LIB=<path>.
set -e
PKG="$(dirname "$0")/../depth-fusion-in-transformer-based-video-object-detection_b200"
name="$1"; src="$2"; shift 2
mkdir -p "$PKG/variants"
( cd "$PKG" && make -j8 >/dev/null )
obj="$PKG/variants/${name}_$(basename "$src" .cu).o"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c "$PKG/csrc/$src" -o "$obj"
others=$(ls "$PKG"/build/*.o | grep -v "/$(basename "$src" .cu).o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$PKG/variants/libmsda_b200_${name}.so" $others "$obj"
echo "$PKG/variants/libmsda_b200_${name}.so"
