#!/bin/bash
# Builds the C++ drop-in check against the header-only layer and librandblas_b200.so (host compiler only).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
LIBDIR="$ROOT/randblas_b200"
g++ -std=c++17 -O2 -Wall -I "$ROOT/include" -I /usr/local/cuda/include -o "$HERE/test_dropin" "$HERE/test_dropin.cc" \
    -L "$LIBDIR" -lrandblas_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$LIBDIR" -Wl,-rpath,/usr/local/cuda/lib64
