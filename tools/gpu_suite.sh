#!/bin/bash
# Run the -m gpu suite file by file (one process each, so a faulted CUDA context cannot cascade), then the kernel
# comparison tool and a short bench.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-run}
for f in test_gpu_kernels test_gpu_network test_gpu_fullsize test_gpu_entrypoints test_gpu_multi; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -x -s > gpurun_out/${tag}_$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/${tag}_summary.log
  tail -3 gpurun_out/${tag}_$f.log
done
timeout 600 python tools/row_vs_tile.py 32 > gpurun_out/${tag}_row_vs_tile.log 2>&1; echo "row_vs_tile exit $?" | tee -a gpurun_out/${tag}_summary.log
cat gpurun_out/${tag}_row_vs_tile.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; echo "bench exit $?" | tee -a gpurun_out/${tag}_summary.log
tail -c 3000 gpurun_out/${tag}_bench.log
