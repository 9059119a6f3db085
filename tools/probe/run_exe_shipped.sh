#!/bin/bash
# d2q9-bgk.exe on the four shipped cases, full iteration counts, with the checker
for c in 128x128 128x256 256x256 1024x1024; do
  d=$(mktemp -d)
  python tools/cases.py $c $d > /dev/null
  echo "== $c"
  (cd $d && /root/repo/hpc-lattice-boltzmann_b200/d2q9-bgk.exe input_$c.params obstacles_$c.dat | grep -E "Reynolds|Elapsed time|GPUs|GPU timestep loop|MLUPS")
  python tools/run_check.py --ref-av-vels-file=tests/golden/$c.npz --ref-final-state-file=tests/golden/$c.npz --av-vels-file=$d/av_vels.dat --final-state-file=$d/final_state.dat 2>&1 | tail -2
  rm -rf $d
done
