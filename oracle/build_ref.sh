#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.  Builds the UNMODIFIED reference (fish_mod.cpp) plus
# the replay harness into oracle/_ref/libfishref_<num>.so, one library per school capacity `num`
# (the reference sizes its arrays with a macro, fish_mod.cpp:23).  The reference source is read
# where it lies under $REF (default /root/reference) and is never copied into the repository:
# the one-line `#define num` edit is made on a temporary file that is deleted after compiling.
# Usage: oracle/build_ref.sh [num ...]      (default: 60 200 1000 10000 1000000)
# REF_OPT / REF_SUFFIX select the optimisation level and a file-name suffix: the parity libraries are -O2; the reference's
# own build script compiles without -O (fishmod_compile.sh:1), so `REF_OPT=-O0 REF_SUFFIX=_O0 build_ref.sh 60 1000` builds
# the as-shipped variants bench.py times for the single-thread CPU baseline.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
out="$here/_ref"
mkdir -p "$out"
[ -f "$REF/fish_mod.cpp" ] || { echo "build_ref: $REF/fish_mod.cpp not found (reference absent: keeping prebuilt $out)"; exit 0; }
nums=("$@"); [ ${#nums[@]} -gt 0 ] || nums=(60 200 1000 10000 1000000)
opt="${REF_OPT:--O2}"; suffix="${REF_SUFFIX:-}"
tmp="$(mktemp -d)"; trap 'rm -rf "$tmp"' EXIT
for n in "${nums[@]}"; do
  # grid stride: the reference hard-codes 250 columns (fish_mod.cpp:24-25); the 1M-fish timing
  # build needs a (W/20)^2 grid, so the stride follows the world size used for that capacity.
  stride=250; [ "$n" -ge 100000 ] && stride=2300
  tu="$tmp/fish_mod_num${n}.cpp"
  sed -e "23s/^#define num 60\$/#define num ${n}/" \
      -e "24s/^#define n_cols 250/#define n_cols ${stride}/" \
      -e "25s/^#define n_rows 250/#define n_rows ${stride}/" "$REF/fish_mod.cpp" > "$tu"
  grep -q "^#define num ${n}\$" "$tu" || { echo "build_ref: line 23 of fish_mod.cpp is not '#define num 60'"; exit 1; }
  g++ -std=c++11 "$opt" -ffp-contract=off -fPIC -shared -w -mcmodel=large -I"$here/shim" -I"$here" -DREF_TU="\"$tu\"" \
      "$here/ref_harness.cpp" -o "$out/libfishref_${n}${suffix}.so" -lpthread
  echo "built $out/libfishref_${n}${suffix}.so"
done
cp "$REF/quality.csv" "$out/quality.csv"
