# Last session of round 2: C2 headline with the straight-line diagonal / padded stages of K2c -- launch list + one
# full capture of k_trmm<false>; likelihood at n = 1024, 8 vectors (eager launches) with the shorter chain.
set -u
mkdir -p gpurun_out
B="--no-cpu-baseline --steps 2 --warmup 3"
python bench.py --no-extra $B > gpurun_out/r02s3_plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s3_c2_launches.csv python bench.py --no-extra $B > gpurun_out/r02s3_ncu_c2.log 2>&1
python tools/agg_launches.py gpurun_out/r02s3_c2_launches.csv > gpurun_out/r02s3_c2_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:k_trmm --launch-skip 12 --launch-count 1 -o /tmp/r02s3_trmm_c2 -f python bench.py --no-extra $B > gpurun_out/ncu_r02s3_trmm_c2.log 2>&1
ncu -i /tmp/r02s3_trmm_c2.ncu-rep --page raw --csv > gpurun_out/r02s3_trmm_c2_raw.csv 2>/dev/null
ncu -i /tmp/r02s3_trmm_c2.ncu-rep --page details > gpurun_out/r02s3_trmm_c2_details.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02s3_trmm_c2_raw.csv > gpurun_out/r02s3_trmm_c2_ncu_full_summary.txt
rm -f /tmp/r02s3_trmm_c2.ncu-rep
python bench.py --workload nll --nll-n 1024 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02s3_plain_nll1024.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02s3_nll1024_launches.csv python bench.py --workload nll --nll-n 1024 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02s3_ncu_nll1024.log 2>&1
python tools/agg_launches.py gpurun_out/r02s3_nll1024_launches.csv > gpurun_out/r02s3_nll1024_launches_summary.txt
cat gpurun_out/r02s3_c2_launches_summary.txt | head -8
cat gpurun_out/r02s3_trmm_c2_ncu_full_summary.txt | head -40
cat gpurun_out/r02s3_nll1024_launches_summary.txt | head -16
