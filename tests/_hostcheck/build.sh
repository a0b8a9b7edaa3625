#!/bin/sh
# TEST-ONLY host build of the shared per-packet physics (see hostcheck.cpp).
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -mfma -x c++ hostcheck.cpp -o libnexo_hostcheck.so
