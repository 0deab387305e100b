# Round-1 evidence capture (run on a B200 box through gpurun): launch lists + `ncu --set full` of the top kernels.
# Exports CSV/text summaries only (the .ncu-rep files exceed the 64 MiB gpurun_out limit); copy what is kept into profiles/.
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
cap() { # name, kernel regex, skip, count, bench args...
  name=$1; shift; rx=$1; shift; sk=$1; shift; cn=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $sk --launch-count $cn -o /tmp/$name -f python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/${name}_source.csv
  rm -f /tmp/$name.ncu-rep
}
cap trmm_c2_r01 k_trmm 12 1 --steps 2 --warmup 3 --no-cpu-baseline
cap rgen_c2_r01 k_rgen 12 1 --steps 2 --warmup 3 --no-cpu-baseline
python bench.py --workload nll --steps 1 --warmup 3 --no-cpu-baseline --nll-groups 1 > gpurun_out/plain_nll.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_nll_r01.csv python bench.py --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1 > gpurun_out/ncu_nll.log 2>&1
cap dgemm_upd_r01 k_dgemm 7 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap dgemm_tri_r01 k_dgemm 71 2 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
ls -la gpurun_out | tail -20; du -sh gpurun_out
