#!/bin/bash
# 4/8-GPU checks of the distributed filter and of the default scaling bench (gpurun --gpus 8)
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 tools/run_dist.py --log2 22 24 26 28 --T 50 2>&1 | grep -E "C5|rror" | tail -6
timeout 600 $TR --master-port 29532 bench.py --gpus $NG --steps 5 --warmup 3 2>/dev/null | grep "^{" | tee gpurun_out/r1_default_x${NG}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default:', d['n_gpus'], d['value']/1e9, d['scaling'], d['e2e']['value']/1e9)"
timeout 600 $TR --master-port 29533 bench.py --gpus $NG --workload c3 --steps 2 --warmup 3 2>/dev/null | grep "^{" | tee gpurun_out/r1_c3_x${NG}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3:', d['n_gpus'], d['value']/1e9, d['scaling'])"
timeout 600 $TR --master-port 29534 bench.py --gpus $NG --workload c5 --steps 3 --warmup 3 2>/dev/null | grep "^{" | tee gpurun_out/r1_c5_x${NG}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5:', d['n_gpus'], d['value']/1e9, d['scaling'], d['ms_per_step'])"
timeout 300 python -m pytest tests/test_gpu_distributed.py -q 2>&1 | tail -1
