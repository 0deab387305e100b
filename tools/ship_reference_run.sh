#!/bin/bash
# One run of the UNMODIFIED reference loop (mipego.BO / ParallelBO, BBOB objectives) on the CUDA
# drop-ins, on a real B200.  The reference is pure Python and lives only in the build container
# (/root/reference); for this one call a copy of the files the run imports is placed under
# baseline/_ref/reference (git-ignored, so it never enters the history; not gpurun-ignored, so it
# travels with the snapshot) and removed again afterwards.  Logs come back in gpurun_out/.
#   tools/ship_reference_run.sh            -> gpurun_out/reference_loop_test.log, gpurun_out/bbob_driver.json
set -u
cd "$(dirname "$0")/.."
REF=${MIPEGO_REFERENCE_ROOT:-/root/reference}
DST=baseline/_ref/reference
rm -rf "$DST"; mkdir -p "$DST/benchmark"
cp -r "$REF/mipego" "$DST/mipego"
cp "$REF/benchmark/bbobbenchmarks.py" "$DST/benchmark/"
find "$DST" -name "__pycache__" -type d -exec rm -rf {} +
trap 'rm -rf baseline/_ref/reference' EXIT
/usr/local/graft/bin/gpurun --timeout ${TIMEOUT:-1500} -- '
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_reference_loop.py -m gpu -q -s -p no:cacheprovider) > gpurun_out/reference_loop_test.log 2>&1
echo "pytest rc=$?" >> gpurun_out/reference_loop_test.log
(time timeout 800 python tools/bbob_driver.py --dim 10 --fids 1,8,15,21 --iters 30 --c5 --out gpurun_out/bbob_driver.json) > gpurun_out/bbob_driver.log 2>&1
echo "driver rc=$?" >> gpurun_out/bbob_driver.log
tail -3 gpurun_out/reference_loop_test.log; tail -3 gpurun_out/bbob_driver.log'
