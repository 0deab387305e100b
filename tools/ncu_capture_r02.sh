# Round-2 evidence capture (run on a B200 box through gpurun): launch lists + `ncu --set full` of the kernels that
# changed this round.  Exports CSV / text only (the .ncu-rep files exceed the gpurun_out limit).  Every command runs
# once WITHOUT ncu first (must exit 0) -- numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --steps 2 --warmup 3"
python bench.py --no-extra $B > gpurun_out/r02_plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c2_launches.csv python bench.py --no-extra $B > gpurun_out/r02_ncu_c2.log 2>&1
python bench.py --workload c2dx $B > gpurun_out/r02_plain_c2dx.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_c2dx_launches.csv python bench.py --workload c2dx $B > gpurun_out/r02_ncu_c2dx.log 2>&1
cap() { # name, kernel regex, skip, count, bench args...
  name=$1; shift; rx=$1; shift; sk=$1; shift; cn=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $sk --launch-count $cn -o /tmp/$name -f python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
cap r02_grad_c2dx k_grad 4 1 --workload c2dx $B
cap r02_trmmfull_c2dx k_trmm 4 1 --workload c2dx $B
cap r02_trmm_c2 k_trmm 12 1 --no-extra $B
# likelihood, n = 4096, 8 vectors, one batch group (serialised launch list) and graphs off (ncu sees every kernel)
python bench.py --workload nll --steps 1 --warmup 3 --no-cpu-baseline --nll-groups 1 > gpurun_out/r02_plain_nll.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/r02_nll_launches.csv python bench.py --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1 > gpurun_out/r02_ncu_nll.log 2>&1
# likelihood at a BO-loop size: n = 1024, 8 vectors
python bench.py --workload nll --nll-n 1024 --steps 1 --warmup 3 --no-cpu-baseline --nll-groups 1 > gpurun_out/r02_plain_nll1024.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_nll1024_launches.csv python bench.py --workload nll --nll-n 1024 --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1 > gpurun_out/r02_ncu_nll1024.log 2>&1
ls -la gpurun_out | tail -30; du -sh gpurun_out
