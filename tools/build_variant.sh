#!/bin/bash
# Build a kernel-variant copy of libdcpgpu.so for A/B timing on the GPU box (selected with DCPGPU_LIB):
#   tools/build_variant.sh <name> "<extra -D flags>"   ->  gpurun_variants/lib_<name>.so
set -e
name=$1; defs=$2
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/gpurun_variants"
make -s -j8 -C "$root/deciphon-old_b200/csrc" BUILD="/tmp/dcp_variant_$name" XDEFS="$defs" OUT="$root/gpurun_variants/lib_$name.so" "$root/gpurun_variants/lib_$name.so"
echo "built gpurun_variants/lib_$name.so ($defs)"
