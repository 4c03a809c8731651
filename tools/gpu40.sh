set -x
mkdir -p gpurun_out
for h in 0 1 2 3; do
echo "NBC_L2_HINTS=$h"
NBC_L2_HINTS=$h timeout 300 python tools/layer_profile.py 8 624 1024 > gpurun_out/layers_hint$h.txt 2>&1; tail -n 1 gpurun_out/layers_hint$h.txt
grep -E "layer3.2.conv1|layer3.2.conv3|layer4.1.conv1|layer4.1.conv3|layer2.2.conv1|layer1.2.conv1" gpurun_out/layers_hint$h.txt
done
