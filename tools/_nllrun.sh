python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in 1024 2048 4096; do python bench.py --workload nll --nll-n $n --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/n.err | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['config']['n'], 'value %.1f evals/s e2e %.1f ms/step %.2f frac %.3f launches %d llf0 %.10f'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac'], j['gpu_launches'], j['config']['llf0']))"; tail -2 gpurun_out/n.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 125 --csv --log-file gpurun_out/launches_nll_s3b.csv python bench.py --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1 > gpurun_out/ncu_nll.log 2>&1
python tools/agg_launches.py gpurun_out/launches_nll_s3b.csv
