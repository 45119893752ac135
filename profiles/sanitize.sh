#!/bin/bash
# compute-sanitizer over smoke-sized runs of every kernel family (run on the GPU box: gpurun -- 'bash profiles/sanitize.sh').
# memcheck: out-of-bounds / misaligned global + shared accesses;  racecheck: shared-memory hazards (the MLP pair kernels'
# activation buffers and mbarriers);  synccheck: invalid __syncwarp / barrier use (the wave kernels' lane-group protocol);
# initcheck: reads of uninitialised device memory (pending-leaf arrays, arenas).  Summaries land in gpurun_out/ and are
# copied to profiles/r2_sanitizer_*.txt.
set -u
OUT=${OUT:-gpurun_out}
mkdir -p "$OUT"
python profiles/sanitize_target.py all > "$OUT/sanitize_plain.log" 2>&1 || { echo "target fails WITHOUT the sanitizer"; tail -5 "$OUT/sanitize_plain.log"; exit 1; }
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  log="$OUT/sanitize_${tool}.log"
  timeout 1200 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 7 python profiles/sanitize_target.py all > "$log" 2>&1
  rc=$?
  echo "== $tool: rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1)  [$(grep -c ' ok' "$log") stages ok]"
done
# Bounds-check build: the device-side index checks of -DBZ_BOUNDS_CHECK (common.cuh BZ_CHECK), for pools where
# compute-sanitizer is closed.  Build the variant where nvcc is (profiles/build_variant.sh bounds -DBZ_BOUNDS_CHECK).
if [ -f build/exp/lib_bounds.so ]; then
  BETAZERO_B200_LIB=$PWD/build/exp/lib_bounds.so python profiles/sanitize_target.py all > "$OUT/sanitize_bounds.log" 2>&1
  echo "== bounds-check build: rc=$?"; grep -E "ok|violations" "$OUT/sanitize_bounds.log"
fi
