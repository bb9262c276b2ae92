mkdir -p /tmp/twl_ds gpurun_out/var
python -c "
import sys; sys.path.insert(0,'.')
from twilight_b200 import synth
synth.make_dataset('rna_100k','/tmp/twl_ds')"
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader,nounits -lms 250 > gpurun_out/var/smi.csv &
SMI=$!
for i in 1 2 3 4 5 6; do
  if [ $((i % 2)) -eq 0 ]; then OPT="wide_workers=0"; else OPT="wide_workers=8"; fi
  rm -rf /tmp/twl_ds/out.aln /tmp/twl_ds/tmp
  s=$(date +%s.%N)
  TWL_OPTIONS=$OPT TWL_STATS=1 build/twilight_b200 -t /tmp/twl_ds/rna_100k.nwk -i /tmp/twl_ds/rna_100k.fa -o /tmp/twl_ds/out.aln -d /tmp/twl_ds/tmp > /dev/null 2> gpurun_out/var/err$i.txt
  echo "run $i $OPT t=$(date +%s) $(grep -o '"device_ms": [0-9.]*' gpurun_out/var/err$i.txt) $(grep -o '"dp_chain": [0-9.]*' gpurun_out/var/err$i.txt) $(grep -o '"level_calls_wall_ms": [0-9.]*' gpurun_out/var/err$i.txt)"
done
kill $SMI
awk -F, '{c[$1]++} END {for (k in c) print "clock", k, c[k]}' gpurun_out/var/smi.csv | sort -k2n | tail -8
sort -t, -k2n gpurun_out/var/smi.csv | tail -2
