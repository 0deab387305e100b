# Final round-2 capture of the C2 headline with the barrier-free K2c epilogue: launch list + one full capture.
set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --steps 2 --warmup 3"
python bench.py --no-extra $B > gpurun_out/r02f_plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_c2_launches.csv python bench.py --no-extra $B > gpurun_out/r02f_ncu_c2.log 2>&1
python tools/agg_launches.py gpurun_out/r02f_c2_launches.csv > gpurun_out/r02f_c2_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:k_trmm --launch-skip 12 --launch-count 1 -o /tmp/r02f_trmm_c2 -f python bench.py --no-extra $B > gpurun_out/ncu_r02f_trmm_c2.log 2>&1
ncu -i /tmp/r02f_trmm_c2.ncu-rep --page raw --csv > gpurun_out/r02f_trmm_c2_raw.csv 2>/dev/null
ncu -i /tmp/r02f_trmm_c2.ncu-rep --page details > gpurun_out/r02f_trmm_c2_details.txt 2>/dev/null
rm -f /tmp/r02f_trmm_c2.ncu-rep
