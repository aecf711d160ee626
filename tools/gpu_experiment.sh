#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/experiment.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== masked variant: compaction on/off"
for nc in 0 1; do
RMG_NO_COMPACT=$nc RMG_BENCH_SPIN_S=0 timeout -k 10 600 python bench.py --steps 6 --warmup 3 --parts masked | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('no_compact=$nc masked', d['masked_variant'])"
done
echo "=== e2e chunk size"
for cv in 66560 262144 1048576; do
RMG_CHUNK_VOXELS=$cv RMG_BENCH_SPIN_S=0 timeout -k 10 600 python bench.py --steps 3 --warmup 3 --parts e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('chunk=$cv e2e', round(e['value']/1e6,2), 'Mvox/s probe', round(e['pinned_h2d_gbs_probe'],1), 'GB/s kernel_ms', round(e['kernel_ms_inside'],2))"
done
echo "=== randomize R=1000"
RMG_BENCH_SPIN_S=0 timeout -k 10 600 python bench.py --steps 3 --warmup 3 --parts randomize | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps(d['randomize']))"
echo "=== done"
