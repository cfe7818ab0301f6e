python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c5 --steps 3 --warmup 3 2>&1 | grep -E "^\{|rror" | cut -c1-1500
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu 2>&1 | grep -E "^\{|rror" | cut -c1-400
python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu 2>&1 | grep -E "^\{|rror" | cut -c1-400
python tools/run_dist.py --log2 24 26 28 --T 20 2>&1 | grep -E "C5|rror" | tail -4
