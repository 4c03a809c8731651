set -x
mkdir -p gpurun_out; rm -f gpurun_out/one_conv_ob.txt
for ob in "" "NBC_OB6=1"; do
for cfg in "256 1024 1 1 1 8 128 128" "64 256 1 1 1 8 256 256" "128 512 1 1 1 8 128 128"; do
  env $ob timeout 120 python tools/prof_one_conv.py $cfg >> gpurun_out/one_conv_ob.txt 2>&1
done
done
cat gpurun_out/one_conv_ob.txt
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -k "pipeline" > gpurun_out/t_pipe.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/t_pipe.log
