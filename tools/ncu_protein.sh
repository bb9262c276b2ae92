#!/bin/bash
# ncu --set full of the two kernels of the protein path (one launch each) on 4096 pairs of 400-aa families
mkdir -p gpurun_out/r02b
for k in simMatrixAaKernel talcoWavefrontKernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02b/prof_protein_$k python tools/dp_ab.py --kind protein --length 400 --seeds 1 --reps 1 > /dev/null 2> gpurun_out/r02b/ncu_$k.err
done
ls -la gpurun_out/r02b
