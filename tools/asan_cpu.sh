#!/bin/bash
# The CPU test suite against an AddressSanitizer build of the library's host code (mesh readers, tile builder,
# partitioning, periodic pairing, control-file front end): builds fvens_b200/variants_asan.so next to the default
# library and runs pytest with it preloaded. libstdc++ is preloaded as well so that the sanitizer finds __cxa_throw.
set -e
cd "$(dirname "$0")/.."
make -C fvens_b200/csrc -j"$(nproc)" EXTRA="-Xcompiler -fsanitize=address,-fno-omit-frame-pointer -g" OBJDIR=build_asan TARGET=../variants_asan.so > /dev/null
ASAN=$(gcc -print-file-name=libasan.so)
STDCXX=$(gcc -print-file-name=libstdc++.so.6)
[ -f "$STDCXX" ] || STDCXX=/usr/lib/x86_64-linux-gnu/libstdc++.so.6
export ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0
LD_PRELOAD="$ASAN $STDCXX" FVENS_B200_LIB=$PWD/fvens_b200/variants_asan.so python -m pytest tests -q -m "not gpu" -p no:cacheprovider "$@"
