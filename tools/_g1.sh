mkdir -p gpurun_out
cd tools
timeout 60 ./diag_bench > ../gpurun_out/diag_v2.txt 2>&1; echo "v2 exit $?" >> ../gpurun_out/diag_v2.txt
cd ..
cat gpurun_out/diag_v2.txt
timeout 300 python -m pytest tests/test_gpu_boundary.py -k "rowwise or graphs" -x -q 2>&1 | tail -5
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 300 python tools/nll_chain_bench.py 500 1024 2048 2>&1 | tee gpurun_out/chain_v2.txt | tail -14
timeout 120 python tools/fit_profile.py 2>&1 | head -3 | tee gpurun_out/fit_v2.txt
