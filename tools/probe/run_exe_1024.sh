#!/bin/bash
d=$(mktemp -d)
python tools/cases.py 1024x1024 $d > /dev/null
cd $d
for v in "" "LBM_GRAPH=0" "LBM_ITERS=40000" "LBM_PDL=0" ""; do
  echo "== 1024x1024 $v"
  env $v /root/repo/hpc-lattice-boltzmann_b200/d2q9-bgk.exe input_1024x1024.params obstacles_1024x1024.dat | grep -E "Elapsed time|GPU timestep loop|MLUPS \(timestep"
done
