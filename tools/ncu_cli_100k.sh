mkdir -p gpurun_out/ncu100k /tmp/twl_ds
python -c "
import sys; sys.path.insert(0,'.')
from twilight_b200 import synth
print(synth.make_dataset('rna_100k','/tmp/twl_ds'))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/ncu100k/launches.csv build/twilight_b200 -t /tmp/twl_ds/rna_100k.nwk -i /tmp/twl_ds/rna_100k.fa -o /tmp/twl_ds/out.aln -d /tmp/twl_ds/tmp > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/ncu100k/launches.csv') if not l.startswith('=='))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
agg=collections.Counter(); n=collections.Counter()
for r in rows[1:]:
    k=r[ik].split('(')[0][:60]; agg[k]+=float(r[iv].replace(',',''))/1e6; n[k]+=1
for k,v in agg.most_common(): print(f"{v:10.2f} ms x{n[k]:5d} {k}")
PY
gzip -f gpurun_out/ncu100k/launches.csv
