#!/usr/bin/env bash
# Builds libdgppo_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles
# without a GPU.  Usage: dgppo_b200/csrc/build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="$root/dgppo_b200/libdgppo_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -I"$root/include" -I"$here" \
  -prec-div=true -prec-sqrt=true "$@" \
  "$here/env_kernels.cu" "$here/reset_kernels.cu" "$here/gnn_kernels.cu" "$here/gnn_v2.cu" "$here/head_tc.cu" "$here/gae_kernels.cu" "$here/rollout.cu" \
  -o "$out" -lcudart
echo "built $out"
