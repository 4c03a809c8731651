set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
for n in 1 2 4; do
timeout 300 python tools/layer_profile.py $n 1024 1024 > gpurun_out/layers_n${n}_1024.txt 2>&1; tail -n 1 gpurun_out/layers_n${n}_1024.txt
done
timeout 300 python tools/layer_profile.py 4 624 1024 > gpurun_out/layers_n4_624.txt 2>&1; tail -n 1 gpurun_out/layers_n4_624.txt
timeout 300 python tools/layer_profile.py 16 624 1024 > gpurun_out/layers_n16_624.txt 2>&1; tail -n 1 gpurun_out/layers_n16_624.txt
