#!/bin/bash
# Multi-GPU round artefacts (run under `gpurun --gpus 8`): kr_group parity + epoch on every device, then the bench at N = 2, 4, 8
# (weak) and N = 8 (strong).  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/group_check.py C5 > gpurun_out/r2_group_c5_n8.json 2> gpurun_out/r2_group_c5_n8.err
python -m pytest tests/test_group.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_group_tests_n8.txt
P=29511
NG=$(python -c "import torch; print(torch.cuda.device_count())")
for N in 2 4 8; do
  [ $N -le $NG ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-pack-leg \
    > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((P+20)) bench.py --gpus $NG --steps 20 --warmup 5 --scaling strong \
  > gpurun_out/r2_bench_n8_strong.json 2> gpurun_out/r2_bench_n8_strong.err
tail -c 600 gpurun_out/r2_group_c5_n8.json gpurun_out/r2_group_tests_n8.txt
for f in gpurun_out/r2_bench_n*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "scaling", "e2e")})
except Exception as e:
    print("unreadable:", e)
PY
done
tail -3 gpurun_out/r2_bench_n8.err
