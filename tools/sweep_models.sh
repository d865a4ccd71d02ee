#!/bin/bash
# Config 3: every fixture model x 100k synthetic sequences, default geometry and the alternatives that fit.
# Usage (on the GPU box): bash tools/sweep_models.sh > gpurun_out/sweep_models.jsonl
cd "$(dirname "$0")/.."
for f in $(ls fixtures/profile_HMMs | sort -n); do
    L=${f%.hmm}
    r4() { echo $(( ($1 + 3) / 4 * 4 )); }
    K8=$(r4 $(( (L + 7) / 8 ))); K16=$(r4 $(( (L + 15) / 16 ))); K32=$(r4 $(( (L + 31) / 32 ))); KW=$(r4 $(( (L + 32) / 32 )))
    geos="default"
    [ $K8 -le 88 ] && geos="$geos 8,$K8"
    [ $K16 -le 88 ] && geos="$geos 16,$K16"
    geos="$geos 32,$K32"
    if [ $KW -ge 24 ]; then geos="$geos 32,$KW,16"; elif [ $KW -ge 8 ]; then geos="$geos 32,$KW,8"; else geos="$geos 32,$KW,0"; fi
    echo "{\"model\": \"$f\"}"
    python tools/sweep_geometry.py --model $f --sequences 100000 --steps 3 --check 64 --geometries $geos 2>&1 | sed -E 's/"chosen": \{"lanes_per_sequence": ([0-9]+), "columns_per_lane": ([0-9]+), "tensor_columns_per_lane": (-?[0-9]+), "threads_per_cta": ([0-9]+), "shared_bytes": ([0-9]+)\}/"G": \1, "K": \2, "KT": \3, "T": \4/'
done
