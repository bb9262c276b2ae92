#!/bin/bash
# repeated CLI runs on the 10^5-leaf set with the host step trace: DP-phase time per run, the kernel time of levels 16-18 and their stage plan
mkdir -p /tmp/twl_ds gpurun_out/var2
python -c "
import sys; sys.path.insert(0,'.')
from twilight_b200 import synth
synth.make_dataset('rna_100k','/tmp/twl_ds')"
for i in 1 2 3 4 5 6; do
  rm -rf /tmp/twl_ds/out.aln /tmp/twl_ds/tmp
  TWL_TRACE=1 TWL_STATS=1 build/twilight_b200 -t /tmp/twl_ds/rna_100k.nwk -i /tmp/twl_ds/rna_100k.fa -o /tmp/twl_ds/out.aln -d /tmp/twl_ds/tmp > /dev/null 2> gpurun_out/var2/err$i.txt
  echo "run $i $(grep -o '"dp_chain": [0-9.]*' gpurun_out/var2/err$i.txt) levels 16-18: $(grep "kernels + D2H" gpurun_out/var2/err$i.txt | awk '{print $(NF-1)}' | sed -n 16,19p | tr '\n' ' ')"
done
for i in 1 2 3 4 5 6; do grep -A1 "dp chain: 154 pairs" gpurun_out/var2/err$i.txt | cut -c1-200; grep -A3 "dp chain: 154 pairs" gpurun_out/var2/err$i.txt | grep "previous dp chain"; done
for i in 1 2 3 4 5 6; do echo "run $i"; grep -A8 "dp chain: 154 pairs" gpurun_out/var2/err$i.txt | grep "launch\|kernels" ; done
