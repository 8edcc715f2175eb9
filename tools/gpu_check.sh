#!/bin/bash
# GPU suite + short benches of both tensor-core precisions (development check)
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for prec in bf16x3 bf16; do
  python bench.py --steps 2 --warmup 2 --quick --no-cpu-baseline --precision $prec 2>/dev/null | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); r=b['roofline']
print('$prec', 'value %.2f e2e %.2f enc %.0f ms dec %.0f ms' % (b['value'], b['e2e']['value'], b['encode_ms_per_gop'], b['decode_ms_per_gop']), 'dominant', r['kernel'][:24], 'frac %.3f' % r['frac'], 'all-tc frac %.3f' % r['all_tensor_stages']['frac'])"
done
nvidia-smi --query-gpu=memory.used --format=csv,noheader
