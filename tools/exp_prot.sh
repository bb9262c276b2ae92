#!/bin/bash
# tools/exp_prot.sh TAG: protein path. Parity tests (DP, level pipeline, CLI --type p), then the DP chain on 4096 pairs of 400-aa families
# with and without the similarity-matrix path (path checksums must agree).
TAG=${1:-prot}; O=gpurun_out/$TAG; mkdir -p $O
timeout 1200 python -m pytest tests/test_protein.py tests/test_cli_synth_gpu.py -x -q -m gpu -k "protein or prot" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log
for ps in 0 1; do
  timeout 600 python tools/dp_ab.py --kind protein --length 400 --seeds 2 --reps 3 --tag $TAG --opts protein_sim=$ps 2> $O/ab$ps.err | python -c "
import json,sys
d=json.load(sys.stdin)
for s in d['seeds']: print('protein_sim=$ps seed', s['seed'], 'ms', round(s['ms_median'],3), 'gcups', round(s['gcups_median'],2), 'launches', s['launches'], 'md5', s['md5'], 'failed', s['failed'])
"
  tail -2 $O/ab$ps.err
done
