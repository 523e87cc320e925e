#!/bin/sh
# builds the CPU emulations of the device code (test infrastructure)
set -e
cd "$(dirname "$0")"
/usr/bin/g++ -O2 -g -std=c++17 -fPIC -shared -ffp-contract=off -I. -o libemu_xdrop.so emu_xdrop.cpp
/usr/bin/g++ -O2 -g -std=c++17 -fPIC -shared -ffp-contract=off -I. -o libemu_seed.so emu_seed.cpp
/usr/bin/g++ -O2 -g -std=c++17 -fPIC -shared -ffp-contract=off -I. -o libemu_map.so emu_map.cpp -L../../oracle -lag2_oracle -Wl,-rpath,'$ORIGIN/../../oracle'
