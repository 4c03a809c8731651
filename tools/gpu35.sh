set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -k "preprocess or trim or engine or pipeline" > gpurun_out/t_pre.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/t_pre.log
PROF="python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"resize4x|trim_rows" -c 200 --csv --log-file gpurun_out/launches_k1.csv $PROF > gpurun_out/ncu35.log 2>&1; echo "ncu exit $?"
grep "resize4x_pass1" gpurun_out/launches_k1.csv | tail -n 3 | cut -d, -f13-
grep "resize4x_pass2" gpurun_out/launches_k1.csv | tail -n 3 | cut -d, -f13-
