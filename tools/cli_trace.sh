#!/bin/bash
# tools/cli_trace.sh DATASET [extra CLI args]: the drop-in CLI on a synthetic data set with the host-side step trace (TWL_TRACE) and the
# device accounting (TWL_STATS); prints the wall-clock breakdown
DS=${1:-rna_100k}; shift
O=gpurun_out/trace_$DS; mkdir -p $O /tmp/twl_ds
python - <<PY
import time, sys
sys.path.insert(0, '.')
from twilight_b200 import synth
t=time.time(); p=synth.make_dataset("$DS", "/tmp/twl_ds"); print("dataset", p, round(time.time()-t,1), "s")
PY
for r in 1 2; do
  rm -rf $O/out.aln $O/tmp
  s=$(date +%s.%N)
  TWL_TRACE=1 TWL_STATS=1 build/twilight_b200 -v -t /tmp/twl_ds/$DS.nwk -i /tmp/twl_ds/$DS.fa -o $O/out.aln -d $O/tmp "$@" > $O/stdout.txt 2> $O/stderr.txt
  e=$(date +%s.%N)
  echo "run $r wall $(echo "$e - $s" | bc) s"
done
grep -h "twl-stats" $O/stderr.txt | cut -c1-600
grep -h "completed in\|Total\|in [0-9.]* s\|seconds" $O/stdout.txt $O/stderr.txt | head -40
python - <<PY
import re, collections
agg=collections.Counter(); n=collections.Counter()
for line in open("$O/stderr.txt"):
    m=re.match(r"\[twl\]\s+(.*?)\s+([0-9.]+) ms", line)
    if m: agg[m.group(1)]+=float(m.group(2)); n[m.group(1)]+=1
for k,v in agg.most_common(25): print(f"{v:10.1f} ms  x{n[k]:5d}  {k}")
PY
md5sum $O/out.aln; rm -f $O/out.aln; rm -rf $O/tmp
