#!/bin/sh
# builds the one-warp CPU emulation of the device code (test infrastructure)
set -e
cd "$(dirname "$0")"
/usr/bin/g++ -O2 -g -std=c++17 -fPIC -shared -I. -o libemu_xdrop.so emu_xdrop.cpp
