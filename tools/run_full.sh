#!/bin/bash
# tools/run_full.sh DATASET [CLI args...]: a BASELINE.json configuration at full size through the drop-in CLI with the host's own --check
# (no golden md5 at these sizes); prints wall clock, the device accounting (TWL_STATS) and the shape of the alignment.
DS=$1; shift
O=gpurun_out/full_$DS; mkdir -p $O /tmp/twl_ds
python - <<PY
import time, sys
sys.path.insert(0, '.')
from twilight_b200 import synth
t=time.time(); p=synth.make_dataset("$DS", "/tmp/twl_ds"); print("dataset", p, round(time.time()-t,1), "s", flush=True)
PY
ls -la /tmp/twl_ds/$DS.*
for r in 1 2; do
  rm -rf /tmp/twl_ds/out.aln /tmp/twl_ds/tmp
  s=$(date +%s%N)
  TWL_STATS=1 timeout 900 build/twilight_b200 -v --check "$@" -t /tmp/twl_ds/$DS.nwk -i /tmp/twl_ds/$DS.fa -o /tmp/twl_ds/out.aln -d /tmp/twl_ds/tmp > $O/stdout.txt 2> $O/stderr.txt
  rc=$?
  e=$(date +%s%N)
  echo "run $r rc=$rc wall $(( (e - s) / 1000000 )) ms"
done
grep -h "twl-stats" $O/stderr.txt | tail -1 | cut -c1-700
grep -h "Total Execution\|Alignment (length" $O/stdout.txt $O/stderr.txt | tail -3
echo "check: $(grep -h 'Completed checking' $O/stderr.txt | awk '{s+=$3} END {print s}') sequences checked, $(grep -c 'did not match' $O/stdout.txt) complaints"
ls -la /tmp/twl_ds/out.aln; md5sum /tmp/twl_ds/out.aln
rm -rf /tmp/twl_ds
