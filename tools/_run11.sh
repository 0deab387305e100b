python tools/latency_check.py 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dist_fit_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -5
