mkdir -p /tmp/twl_ds gpurun_out/var3
python -c "
import sys; sys.path.insert(0,'.')
from twilight_b200 import synth
synth.make_dataset('rna_100k','/tmp/twl_ds')"
for i in 1 2 3 4 5 6; do
  rm -rf /tmp/twl_ds/out.aln /tmp/twl_ds/tmp
  TWL_OPTIONS=dp_trace=1 TWL_TRACE=1 TWL_STATS=1 build/twilight_b200 -t /tmp/twl_ds/rna_100k.nwk -i /tmp/twl_ds/rna_100k.fa -o /tmp/twl_ds/out.aln -d /tmp/twl_ds/tmp > /dev/null 2> gpurun_out/var3/err$i.txt
  echo "run $i $(grep -o '"dp_chain": [0-9.]*' gpurun_out/var3/err$i.txt)"
done
python - <<PY
import re
for i in range(1,7):
    lv=-1; out=[]
    for line in open(f"gpurun_out/var3/err{i}.txt"):
        if "describe level" in line: lv+=1
        if line.startswith("[twl dp]") and 15<=lv<=18: out.append(f"L{lv} "+line.strip()[9:])
    print("run",i); print("\n".join(out))
PY
