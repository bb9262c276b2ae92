#!/bin/bash
# One GPU call of the DP-kernel development loop: parity of the DP kernels, throughput A/B (path checksums), lone-pair latency,
# and the drop-in CLI on the two bundled data sets (device time from TWL_STATS, md5 of the FASTA).
#   tools/exp.sh TAG [extra dp_ab opts]
TAG=${1:-exp}; OPTS=${2:-}
O=gpurun_out/$TAG; mkdir -p $O
D=oracle/_ref/dataset
timeout 900 python -m pytest tests/test_talco_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" ; tail -3 $O/tests.log
timeout 600 python tools/dp_ab.py --seeds 2 --reps 3 --tag $TAG ${OPTS:+--opts $OPTS} > $O/ab.json 2> $O/ab.err; cat $O/ab.json | python -c "
import json,sys
d=json.load(sys.stdin)
for s in d['seeds']: print('dp_ab seed', s['seed'], 'ms', round(s['ms_median'],2), 'gcups', round(s['gcups_median'],1), 'md5', s['md5'], 'failed', s['failed'])
"
for m in 1 64; do timeout 300 python tools/one_pair.py 1 $m 2>&1 | tail -1; done
timeout 300 python tools/one_pair.py 64 8 2>&1 | tail -1
for ds in sars_20 RNASim; do
  for r in 1 2; do
    rm -rf $O/tmp_$ds $O/$ds.aln
    TWL_STATS=1 timeout 300 build/twilight_b200 -t $D/$ds.nwk -i $D/$ds.fa -o $O/$ds.aln -d $O/tmp_$ds > /dev/null 2> $O/$ds.err
  done
  echo "$ds md5 $(md5sum < $O/$ds.aln | cut -c1-32) $(grep -o '"device_ms": [0-9.]*' $O/$ds.err) $(grep -o '"dp_chain": [0-9.]*' $O/$ds.err) $(grep -o '"level_calls_wall_ms": [0-9.]*' $O/$ds.err)"
  rm -rf $O/tmp_$ds $O/$ds.aln
done
