#!/bin/bash
# tools/run_1m.sh [M] [DEVICES]: BASELINE.json's C3 at full size — 10^6 synthetic RNA leaves x 1.5 kb through the drop-in CLI in
# divide-and-conquer mode (-m M), alignment checked by the host's own --check; prints wall clock, device accounting and sizes.
M=${1:-50000}; DEV=${2:-0}
O=gpurun_out/run_1m; mkdir -p $O /tmp/twl_ds
python - <<PY
import time, sys
sys.path.insert(0, '.')
from twilight_b200 import synth
t=time.time(); p=synth.make_dataset("rna_1m", "/tmp/twl_ds"); print("dataset", p, round(time.time()-t,1), "s", flush=True)
PY
ls -la /tmp/twl_ds/rna_1m.*; free -g | head -2
s=$(date +%s)
TWL_DEVICES=$DEV TWL_STATS=1 timeout 800 build/twilight_b200 -v --check -m $M -t /tmp/twl_ds/rna_1m.nwk -i /tmp/twl_ds/rna_1m.fa -o /tmp/twl_ds/out_1m.aln -d /tmp/twl_ds/tmp_1m > $O/stdout.txt 2> $O/stderr.txt
rc=$?
e=$(date +%s)
echo "rc=$rc wall $((e - s)) s"
grep -h "twl-stats" $O/stderr.txt | tail -3 | cut -c1-700
grep -ch "twl-stats" $O/stderr.txt
grep -h "completed in\|Total\|Finished\|ERROR\|legal\|error" $O/stdout.txt $O/stderr.txt | tail -15
ls -la /tmp/twl_ds/out_1m.aln; head -c 200 /tmp/twl_ds/out_1m.aln | head -2 | cut -c1-80
python - <<PY
n=0; L=set()
with open("/tmp/twl_ds/out_1m.aln","rb") as f:
    cur=0
    for line in f:
        if line.startswith(b">"):
            if n: L.add(cur)
            n+=1; cur=0
        else: cur+=len(line)-1
    L.add(cur)
print("rows", n, "row lengths", sorted(L)[:5])
PY
free -g | head -2; df -h /tmp | tail -1
rm -rf /tmp/twl_ds
