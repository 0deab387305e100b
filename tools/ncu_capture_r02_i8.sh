# Round-2 evidence for the INT8 (Ozaki-scheme) K2c: launch lists of the c2_i8 / c3_i8 workloads and one
# `ncu --set full` capture of k_trmm_i8 and of K1 with the fused slicing.  Every command runs once WITHOUT ncu
# first (must exit 0) -- numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --steps 2 --warmup 3"
for w in c2_i8 c3_i8; do
  python bench.py --workload $w $B > gpurun_out/r02_plain_$w.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_${w}_launches.csv python bench.py --workload $w --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r02_ncu_$w.log 2>&1
  python tools/agg_launches.py gpurun_out/r02_${w}_launches.csv > gpurun_out/r02_${w}_launches_summary.txt
done
cap() { # name, kernel regex, skip, count, bench args...
  name=$1; shift; rx=$1; shift; sk=$1; shift; cn=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $sk --launch-count $cn -o /tmp/$name -f python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
cap r02_trmm_i8_c3 k_trmm_i8 6 1 --workload c3_i8 --no-cpu-baseline --steps 1 --warmup 3
cap r02_rgen_i8_c2 k_rgen 6 1 --workload c2_i8 --no-cpu-baseline --steps 1 --warmup 3
cap r02_trmm_i8_c2 k_trmm_i8 6 1 --workload c2_i8 --no-cpu-baseline --steps 1 --warmup 3
ls -la gpurun_out | tail -20; du -sh gpurun_out
