#!/bin/bash
# gpurun --gpus 2 call: the whole GPU suite on a 2-GPU box (exercises every n_gpus = 2 path), smoke, the 2-rank bench
# (weak scaling of the fit, strong scaling of mincRandomize incl. its end-to-end legs), and the pageable-host legs.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/two.log) 2>&1
nvidia-smi --query-gpu=name,memory.total --format=csv
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu (2 GPUs visible)"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
export RMG_BENCH_SPIN_S=0
echo "=== bench N=2"
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 5 --warmup 3 --parts randomize,stored,e2e,e2e_u16 > gpurun_out/bench_n2_new.json 2> gpurun_out/bench_n2_new.err
echo "rc=$?"; tail -3 gpurun_out/bench_n2_new.err
echo "=== bench N=1: tfce + pageable legs"
timeout -k 10 1200 python bench.py --steps 3 --warmup 3 --parts tfce,e2e_pageable > gpurun_out/bench_tfce_pageable.json 2> gpurun_out/bench_tfce_pageable.err
echo "rc=$?"; tail -3 gpurun_out/bench_tfce_pageable.err
python - <<'PY'
import json
for f, keys in (("gpurun_out/bench_n2_new.json", ("value", "randomize", "e2e")), ("gpurun_out/bench_tfce_pageable.json", ("vertexTFCE", "e2e"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, json.dumps({k: d.get(k) for k in keys}, indent=1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "=== done"
