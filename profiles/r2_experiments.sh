#!/bin/bash
# A/B timing of tree-kernel build variants at the headline configuration (4096 games, 800 sims, 4 leaves): each variant is
# a separate library under build/exp/ (built in the authoring container with the -D flags named in profiles/README.md) and
# is selected with BETAZERO_B200_LIB.  Prints us per MCTS iteration (tree kernel + MLP kernel, PDL chained, CUDA graph).
OUT=${OUT:-gpurun_out}
for lib in ${VARIANTS:-base}; do
  echo "== $lib"
  BETAZERO_B200_LIB=$PWD/build/exp/lib_$lib.so LEAVES=4 python profiles/leaves_probe.py 2>&1 | tail -1
done
