"""Where a predict step's time goes (run on the GPU box): the same 64 device-resident scans through
   (a) the network passes alone (ragged chunks back to back), (b) K1 alone, (c) K3 + K5 alone, (d) the whole engine.
usage: python tools/step_breakdown.py [n_images]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import neuralbarkcalculator_b200 as nbc  # noqa: E402
from neuralbarkcalculator_b200 import engine, ops  # noqa: E402
from oracle import model as omodel  # noqa: E402


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device('cuda:0')
    g = np.load(os.path.join(bench.ROOT, 'tests', 'golden', 'model_small.npz'))
    sd = omodel.synthetic_state_dict(seed=0, head=(g['head_w'], g['head_b']))
    calc = nbc.NeuralBarkCalculator(None, 'cuda:0', state_dict=sd)
    eng = engine.PredictEngine(calc.model, dev)
    raws, _ = bench.synth_raw_gpu(n, dev, 0)
    counts, masks, heights = eng.run_device(raws, exclude_nodes=True)
    torch.cuda.synchronize()
    slot = eng._slots[0] if eng._slots[0] is not None else eng._slots[1]
    plan = calc.model.native_plan()
    C = eng.chunk
    chunks = [(a, min(a + C, n)) for a in range(0, n, C)]
    logits = [torch.empty_like(eng._logits[0]) for _ in chunks]

    def net():
        for k, (a, b) in enumerate(chunks):
            plan.forward_ragged(slot.proc[a:b], heights=slot.heights[a:b], out=logits[k][:b - a])

    def k1():
        for (a, b) in chunks:
            eng._preprocess_chunk(slot, a, b, [raws[i].data_ptr() for i in range(a, b)], [(0, eng.raw_size)] * (b - a), True, True)
            ops.heights_from_first_last(slot.fl[a:b], out=slot.heights[a:b])

    def k3():
        for k, (a, b) in enumerate(chunks):
            ops.upsample_argmax_ragged(logits[k][:b - a], slot.heights[a:b], (1024, 1024), out=slot.masks[a:b])

    def k5():
        for k, (a, b) in enumerate(chunks):
            ops.remove_small_zones_ragged(slot.masks[a:b], slot.heights[a:b], 150, True, workspace=eng._ccl_ws, counts=slot.counts[a:b])

    def whole():
        eng.run_device(raws, exclude_nodes=True)

    t = {'network': timeit(net), 'K1': timeit(k1), 'K3': timeit(k3), 'K5': timeit(k5), 'whole engine step': timeit(whole)}
    print('%d scans, chunk %d, mean trimmed rows %.0f' % (n, C, float(heights.float().mean())))
    for k, v in t.items():
        print('%-20s %8.3f ms  (%.3f ms per chunk of %d, %.1f us per image)' % (k, v, v / len(chunks), C, 1000 * v / n))
    s = t['network'] + t['K1'] + t['K3'] + t['K5']
    print('sum of parts %.3f ms vs whole %.3f ms -> %.1f %% of the step is not kernel time of these four (gaps, cross-stream '
          'waits, memsets)' % (s, t['whole engine step'], 100 * (t['whole engine step'] - s) / t['whole engine step']))


if __name__ == '__main__':
    main()
