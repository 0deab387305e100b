for w in c2 c3 nll; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --workload $w --no-cpu-baseline 2>gpurun_out/b8.err > gpurun_out/bench_${w}_n8_r01b.json; tail -2 gpurun_out/b8.err | cut -c1-300; grep '^{' gpurun_out/bench_${w}_n8_r01b.json | python -c "
import sys,json
j=json.loads(sys.stdin.read())
print('$w n_gpus', j['n_gpus'], 'value %.4g e2e %.4g ms/step %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step']), j['clocks'])"
done
