ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_nll_s3a.csv python bench.py --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1 > gpurun_out/ncu_nll.log 2>&1
python tools/agg_launches.py gpurun_out/launches_nll_s3a.csv
