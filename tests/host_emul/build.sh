#!/bin/sh
# Builds the test-only CPU replay of the kernel bodies (see emul.cpp).
set -e
cd "$(dirname "$0")"
g++ -O2 -fopenmp -std=c++17 -x c++ -shared -fPIC -o liblsted_emul.so emul.cpp -lm
