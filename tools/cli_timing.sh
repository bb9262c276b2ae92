#!/bin/bash
# times the drop-in CLI (both pipelines) against the reference CLI on the bundled datasets; run on the GPU box
D=oracle/_ref/dataset; O=${1:-/tmp/twl_cli}; rm -rf $O; mkdir -p $O
t() { local s=$(date +%s%N); "$@" > /dev/null 2> $O/err.txt; local e=$(date +%s%N); echo "$(( (e - s) / 1000000 )) ms"; }
for ds in sars_20 RNASim; do
  echo "$ds cpu-ref ($(nproc) threads): $(t oracle/_ref/twilight_ref -t $D/$ds.nwk -i $D/$ds.fa -o $O/${ds}_ref.aln)"
  grep -h "completed in\|Total Exec" $O/err.txt | tr '\n' ';'; echo
  echo "$ds b200 level pipeline (1st run of the process pays CUDA context creation): $(t env TWL_PIPELINE=level build/twilight_b200 -v -t $D/$ds.nwk -i $D/$ds.fa -o $O/${ds}_level.aln)"
  grep -h "aligned\|completed in\|Total Exec" $O/err.txt | tr '\n' ';' | cut -c1-900; echo
  echo "$ds b200 dp-only pipeline: $(t env TWL_PIPELINE=dp build/twilight_b200 -v -t $D/$ds.nwk -i $D/$ds.fa -o $O/${ds}_dp.aln)"
  grep -h "aligned\|completed in\|Total Exec" $O/err.txt | tr '\n' ';' | cut -c1-900; echo
done
md5sum $O/*.aln
