set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket'
python -c "import botorch" 2>&1 | tail -1
python -c "import gpytorch" 2>&1 | tail -1
./tools/microbench/fp64_peak > gpurun_out/fp64_peak.txt 2>&1; cat gpurun_out/fp64_peak.txt
python tools/microbench/dgemm_peak.py > gpurun_out/dgemm_peak.json 2>gpurun_out/dgemm_err.txt; cat gpurun_out/dgemm_peak.json
