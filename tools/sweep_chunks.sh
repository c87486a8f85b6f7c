#!/bin/bash
# chunk-size sweep of the fused scorer (L2-resident feature scratch); run on the GPU box
for c in "$@"; do
  WW_CHUNK_CLIPS=$c python bench.py --no-cpu --cnn tensor --steps 4 --e2e-clips 65536 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk', $c, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))"
done
