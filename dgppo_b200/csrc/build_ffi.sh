#!/usr/bin/env bash
# Builds libdgppo_ffi.so (XLA-FFI handlers over libdgppo_b200.so) where jaxlib's headers exist; otherwise reports
# that and exits 0 (this image has no jax: the C ABI + the ctypes mirror are the boundary that is exercised here).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
inc="$(python -c 'import jax.ffi; print(jax.ffi.include_dir())' 2>/dev/null || true)"
if [ -z "$inc" ]; then
  echo "build_ffi: jax.ffi not importable - skipping libdgppo_ffi.so (dgppo_ffi.cc is an empty translation unit without xla/ffi/api/ffi.h)"
  # still prove the file is syntactically inert without the headers
  g++ -std=c++17 -fsyntax-only -I"$root/include" "$here/dgppo_ffi.cc"
  exit 0
fi
g++ -std=c++17 -O2 -shared -fPIC -I"$inc" -I/usr/local/cuda/include -I"$root/include" "$here/dgppo_ffi.cc" \
  -L"$root/dgppo_b200" -ldgppo_b200 -Wl,-rpath,'$ORIGIN' -o "$root/dgppo_b200/libdgppo_ffi.so"
echo "built $root/dgppo_b200/libdgppo_ffi.so"
