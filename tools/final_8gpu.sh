#!/bin/bash
# End-of-round multi-GPU set on one 8-GPU box: the bench under torchrun (one process per GPU) and the drop-in CLI with all GPUs in one process
O=gpurun_out/r02c; mkdir -p $O /tmp/twl_ds
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 6 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_8gpu.json"))
print({k:d[k] for k in ("metric","value","n_gpus","ms_per_step","scaling")}, "e2e", round(d["e2e"]["value"],1), "msa_sharded", (d.get("msa_sharded") or {}).get("seqs_per_s_e2e"))
PY
python -c "
import sys; sys.path.insert(0,'.')
from twilight_b200 import synth
synth.make_dataset('rna_100k','/tmp/twl_ds')"
for dev in 0 all; do
  for r in 1 2; do
    rm -rf /tmp/twl_ds/out.aln /tmp/twl_ds/tmp
    s=$(date +%s%N)
    TWL_DEVICES=$dev TWL_STATS=1 build/twilight_b200 -t /tmp/twl_ds/rna_100k.nwk -i /tmp/twl_ds/rna_100k.fa -o /tmp/twl_ds/out.aln -d /tmp/twl_ds/tmp > /dev/null 2> $O/cli8_$dev.err
    e=$(date +%s%N)
  done
  echo "rna_100k TWL_DEVICES=$dev wall $(( (e - s) / 1000000 )) ms md5 $(md5sum < /tmp/twl_ds/out.aln | cut -c1-32) $(grep -h 'twl-stats' $O/cli8_$dev.err)"
done | tee $O/multigpu_cli_8gpu.txt
rm -rf /tmp/twl_ds
