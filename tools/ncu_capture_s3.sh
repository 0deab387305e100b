# Round-1 (third session) evidence refresh after the likelihood-path changes: final bench lines, the
# NLL launch list (one batch group = serialised) and `ncu --set full` of the changed GEMM / diag kernels.
set -x
for w in c1 c2 c3 c5; do python bench.py --workload $w --steps 6 --warmup 3 > gpurun_out/s3_${w}.json 2> gpurun_out/s3_${w}.err; done
python bench.py --workload nll --steps 8 --warmup 3 > gpurun_out/s3_nll.json 2> gpurun_out/s3_nll.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3_ref_c2.json 2> gpurun_out/s3_ref.err
python bench.py --workload nll --steps 1 --warmup 3 --no-cpu-baseline --nll-groups 1 > gpurun_out/plain_nll_s3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_nll_s3.csv python bench.py --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1 > gpurun_out/ncu_nll_s3.log 2>&1
python tools/agg_launches.py gpurun_out/launches_nll_s3.csv > gpurun_out/launches_nll_s3_summary.txt
cap() { # name, kernel regex, skip, count, bench args...
  name=$1; shift; rx=$1; shift; sk=$1; shift; cn=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $sk --launch-count $cn -o /tmp/$name -f python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
# k_dgemm launches of one evaluation batch (G = 1): 66 x <0,0>, then 10 x <0,1> (trtri), then <1,1> (M^T M)
cap dgemm_lauum_s3 'k_dgemm<1' 0 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap dgemm_trtri_s3 'k_dgemm<0, 1' 9 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap potrf_diag_s3 k_potrf_diag 3 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap buildR_s3 k_build_R 0 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
cap nllgrad_s3 k_nll_grad 0 1 --workload nll --steps 1 --warmup 0 --no-cpu-baseline --nll-groups 1
for f in dgemm_lauum_s3 dgemm_trtri_s3 potrf_diag_s3 buildR_s3 nllgrad_s3; do python tools/ncu_summary.py gpurun_out/${f}_raw.csv > gpurun_out/${f}_summary.txt; done
ls -la gpurun_out | tail -30; du -sh gpurun_out
