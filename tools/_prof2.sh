for g in 0; do python bench.py --workload nll --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('value %.1f evals/s e2e %.1f ms/step %.2f frac %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac']))"; done
cap() { name=$1; shift; rx=$1; shift; sk=$1; shift; cn=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $sk --launch-count $cn -o /tmp/$name -f python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/${name}_source.csv
  rm -f /tmp/$name.ncu-rep; }
cap nllgrad_r01 k_nll_grad 0 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap buildR_r01 k_build_R 0 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
