# Round-1 (third session) final evidence: bench lines, launch lists of the C2 and likelihood workloads
# (likelihood with one batch group = serialised), `ncu --set full` of the kernels changed in this session.
# Build the micro-benchmarks first (in the build container; the binaries travel with the snapshot):
#   cd tools && for t in diag_bench fp64_lat exp_check; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../mip-ego_b200/csrc -I../include -o $t $t.cu; done
# Exports CSV/text summaries only (the .ncu-rep files exceed the 64 MiB gpurun_out limit); copy what is kept into profiles/.
for w in c1 c2 c3 c5; do python bench.py --workload $w --steps 6 --warmup 3 > gpurun_out/s3_${w}.json 2> gpurun_out/s3_${w}.err; done
python bench.py --workload nll --steps 8 --warmup 3 > gpurun_out/s3_nll.json 2> gpurun_out/s3_nll.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3_ref_c2.json 2> gpurun_out/s3_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_s3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_s3.log 2>&1
python tools/agg_launches.py gpurun_out/launches_c2_s3.csv > gpurun_out/launches_c2_s3_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_nll_s3.csv python bench.py --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1 > gpurun_out/ncu_nll_s3.log 2>&1
python tools/agg_launches.py gpurun_out/launches_nll_s3.csv > gpurun_out/launches_nll_s3_summary.txt
cap() { # name, kernel regex, skip, count, bench args...
  name=$1; shift; rx=$1; shift; sk=$1; shift; cn=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $sk --launch-count $cn -o /tmp/$name -f python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${name}_raw.csv > gpurun_out/${name}_summary.txt
  rm -f /tmp/$name.ncu-rep
}
cap rgen_c2_s3 k_rgen 12 1 --steps 2 --warmup 3 --no-cpu-baseline
cap trmm_c2_s3 k_trmm 12 1 --steps 2 --warmup 3 --no-cpu-baseline
# k_dgemm launches of one evaluation batch (G = 1): #7 = first k = 512 trailing update, #70/#71 = top-level TRTRI, #72 = M^T M
cap dgemm_tri_s3 k_dgemm 70 3 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap dgemm_upd_s3 k_dgemm 7 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap potrf_diag_s3 k_potrf_diag 3 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap buildR_s3 k_build_R 0 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap nllgrad_s3 k_nll_grad 0 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
timeout 60 tools/diag_bench > gpurun_out/diag_bench_s3.txt 2>&1
tools/fp64_lat > gpurun_out/fp64_lat_s3.json 2>&1
tools/exp_check > gpurun_out/exp_check_s3.json 2>&1
du -sh gpurun_out
