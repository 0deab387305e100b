run() { python bench.py --workload nll --nll-n $1 --steps 4 --warmup 2 --no-cpu-baseline --nll-groups $2 2>gpurun_out/n.err | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$3 n=$1 G=$2', 'value %.1f ms/step %.2f frac %.3f launches %d'%(j['value'], j['ms_per_step'], j['roofline']['frac'], j['gpu_launches']))"; }
for sch in 0 1; do for pr in "" 1; do for g in 4 8; do
  export MGP_CHOL_SCHEME=$sch; if [ -n "$pr" ]; then export MGP_NLL_PRIO=1; else unset MGP_NLL_PRIO; fi
  run 4096 $g "scheme=$sch prio=$pr"
done; done; done
unset MGP_NLL_PRIO
for sch in 0 1; do export MGP_CHOL_SCHEME=$sch; for n in 1024 2048; do for g in 1 2 4; do run $n $g "scheme=$sch"; done; done; done
