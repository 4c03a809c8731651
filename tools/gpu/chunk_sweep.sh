# network pass per image as a function of the chunk size (L2 residency of the block outputs vs tiles per launch)
set -x
mkdir -p gpurun_out
for n in 2 3 4 6 8 12; do
  timeout 200 python tools/layer_profile.py $n 624 1024 > gpurun_out/layers_chunk$n.txt 2>&1
  echo "chunk $n"; tail -n 1 gpurun_out/layers_chunk$n.txt
done
