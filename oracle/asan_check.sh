#!/bin/sh
# One AddressSanitizer + UBSan pass of the C restatement (TEST INFRASTRUCTURE): SURVEY.md section 5 lists undefined
# behaviour on the reference's live path (rescue_clipped_align reads alnv[0] when naln == 0, aux.cpp:428; uninitialised
# maxi, impl_large.cpp:611); this checks that the restatement, which follows those lines, stays inside its buffers on the
# stress fixture and still writes the reference binary's thread file.  Output: profiles/asan_oracle_<round>.txt
set -e
cd "$(dirname "$0")/.."
OUT=${1:-profiles/asan_oracle_r02.txt}
SO=/tmp/libag2_oracle_asan.so
FLAGS="-O1 -g -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -Wno-unused-function"
/usr/bin/gcc $FLAGS -std=gnu11 -c -o /tmp/asan_o1.o oracle/ag2_oracle.c
/usr/bin/gcc $FLAGS -std=gnu11 -c -o /tmp/asan_o2.o oracle/ag2_mapper.c
/usr/bin/gcc $FLAGS -std=gnu11 -c -o /tmp/asan_o3.o oracle/ag2_kmer.c
/usr/bin/g++ $FLAGS -std=c++14 -c -o /tmp/asan_o4.o oracle/ag2_pagraph.cpp
/usr/bin/gcc $FLAGS -std=gnu11 -c -o /tmp/asan_o5.o oracle/ag2_diff.c
/usr/bin/g++ -shared -fsanitize=address,undefined -o $SO /tmp/asan_o1.o /tmp/asan_o2.o /tmp/asan_o3.o /tmp/asan_o4.o /tmp/asan_o5.o -lm
{
  echo "# ASan + UBSan pass of the oracle (oracle/asan_check.sh), $(date -u +%Y-%m-%dT%H:%MZ)"
  echo "# tests: the oracle against the committed goldens of the reference (thread file of the stress fixture, X-drop blocks and"
  echo "# extensions, A-Bruijn dump) and, in-process, against libref_mecat.so; tests that spawn the uninstrumented reference"
  echo "# binaries are left out (they would inherit the preloaded sanitizer runtime)"
  LD_PRELOAD="$(/usr/bin/gcc -print-file-name=libasan.so) $(/usr/bin/gcc -print-file-name=libubsan.so)" \
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 AG2_ORACLE_SO=$SO \
    python -m pytest tests/test_oracle_mapper.py tests/test_oracle_pinned.py tests/test_oracle_pagraph.py tests/test_kmer_counter.py tests/test_oracle_diff.py -q \
      -k "golden or edge_cases or oracle_vs_reference or committed or against_the_reference" 2>&1 | grep -v "^$" | tail -40
} > "$OUT" 2>&1
tail -5 "$OUT"
