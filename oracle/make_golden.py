"""Pins the oracle against the REFERENCE'S OWN CODE and writes the golden fixtures under ``tests/golden/``.

Run in the build container only (needs ``/root/reference``; the GPU box never runs this):

    python -m oracle.make_golden

The reference (``/root/reference/src/bark_calculator``) cannot be imported as is: scikit-image, matplotlib,
efficientnet_pytorch and poutyne are not installed.  None of those is needed by the functions pinned here, so
they are replaced by inert stub modules and the reference modules are then imported unmodified:

* ``models.fcn_resnet50`` / ``SimpleSegmentationModel`` / ``FCNHead``  -> logits, argmax      (a5-a8)
* ``models.trim_black``                                               -> trim rows           (a4)
* ``utils.CustomWeightedCrossEntropy``                                -> loss + gradient     (a13)
* ``dataset.make_dataset_for_dir`` / ``RegressionDatasetFolder``      -> enumeration order   (a2)
* ``predict.generate_folders``                                        -> folder layout       (a1)

Each is compared with the oracle restatement (must agree exactly, or to float round-off where noted) and the
agreed values are stored as fixtures.  What is NOT pinnable this way -- the skimage arithmetic of resize, imsave
and remove_small_holes/objects -- is generated from the oracle alone and flagged ``unpinned`` in the fixture.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = '/root/reference/src/bark_calculator'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    def _missing(*a, **k):
        raise RuntimeError('stubbed third-party function called')

    _stub('skimage')
    _stub('skimage.transform', resize=_missing)
    _stub('skimage.io', imsave=_missing, imread=_missing)
    _stub('skimage.morphology', remove_small_objects=_missing, remove_small_holes=_missing)
    _stub('skimage.segmentation', find_boundaries=_missing)
    _stub('efficientnet_pytorch', EfficientNet=object)
    mpl = _stub('matplotlib', cm=None)
    _stub('matplotlib.pyplot')
    _stub('matplotlib.patches')
    mpl.pyplot = sys.modules['matplotlib.pyplot']
    mpl.patches = sys.modules['matplotlib.patches']
    _stub('poutyne')
    _stub('poutyne.framework')
    _stub('poutyne.framework.callbacks', Callback=object)
    sys.path.insert(0, REF)
    import dataset as ref_dataset
    import models as ref_models
    import predict as ref_predict
    import utils as ref_utils
    return ref_models, ref_utils, ref_dataset, ref_predict


def main():
    import warnings
    warnings.filterwarnings('ignore')
    from oracle import losses as olosses
    from oracle import model as omodel
    from oracle import pipeline as opipe
    from oracle import postprocess as opost
    from oracle import preprocess as opre
    from oracle import synth

    ref_models, ref_utils, ref_dataset, ref_predict = import_reference()
    os.makedirs(OUT, exist_ok=True)
    report = []

    # ---- model: reference fcn_resnet50 vs oracle, same state_dict, small input ------------------------------
    sd = omodel.synthetic_state_dict(seed=0)
    ref_net = ref_models.fcn_resnet50(pretrained=False)
    missing = ref_net.load_state_dict(sd, strict=True)
    ref_net.eval()
    ora_net = omodel.load_model(sd)
    assert list(ref_net.state_dict().keys()) == list(ora_net.state_dict().keys()) and len(sd) == 326
    img = synth.texture_u8(96, 160, seed=11)
    x = omodel.normalise_u8(img)
    with torch.no_grad():
        ref_logits = ref_net(x)
        ora_logits = ora_net(x)
        ora_low = omodel.lowres_logits(ora_net, x)
    assert torch.equal(ref_logits, ora_logits), 'oracle model != reference model'
    ref_mask = torch.argmax(ref_logits, dim=1)
    report.append('model: oracle logits bit-identical to reference models.fcn_resnet50 (eval), 96x160 input')
    np.savez_compressed(os.path.join(OUT, 'model_small.npz'), image=img, lowres_logits=ora_low.numpy(),
                        logits=ref_logits.numpy().astype(np.float32), mask=ref_mask.numpy().astype(np.uint8),
                        state_dict_seed=np.int64(0), head_w=sd['classifier.4.weight'].numpy(),
                        head_b=sd['classifier.4.bias'].numpy())

    # ---- upsample + argmax on its own (reference does it inside forward) --------------------------------------
    g = torch.Generator().manual_seed(5)
    low = torch.randn(2, 3, 13, 16, generator=g)
    up, am = omodel.upsample_argmax(low, (100, 128))
    np.savez_compressed(os.path.join(OUT, 'upsample_argmax.npz'), lowres=low.numpy(), up=up.numpy(),
                        mask=am.numpy().astype(np.uint8))

    # ---- trim_black: reference function on the float image the reference would see ----------------------------
    raw, top, bottom = synth.raw_image_u8(seed=7, size=512, top=77, bottom=130, specks=40)
    raw[200, :100] = 0              # a row with ~20 % dark pixels: must be kept (85 % rule)
    raw[300, :] = 0                 # a fully dark row inside: kept (only leading/trailing rows are cut)
    S = opre.resize4x_S(raw)
    lo, hi = int(raw.min()), int(raw.max())
    f = np.clip(S, 256 * lo, 256 * hi).astype(np.float64) / (256.0 * 255.0)
    ref_trim = ref_models.trim_black(f)
    first = int(np.argmax(np.mean(np.sum(f, -1) > 1e-3, -1) > 0.85))
    _, o_first, o_last = opre.trim_black_float(f)
    assert ref_trim.shape[0] == o_last - o_first and o_first == first
    # integer rule used by the oracle / the CUDA kernel must give the same rows
    Sc = np.clip(S, 256 * lo, 256 * hi)
    i_first, i_last = opre.trim_rows_from_counts((Sc.sum(-1) >= 66).sum(1), Sc.shape[1], Sc.shape[0])
    assert (i_first, i_last) == (o_first, o_last)
    report.append('trim_black: reference rows == oracle float rule == integer rule (%d, %d)' % (o_first, o_last))
    out_u8 = ((Sc + 128) >> 8).astype(np.uint8)[o_first:o_last]
    np.savez_compressed(os.path.join(OUT, 'preprocess_small.npz'), raw=raw, out=out_u8, first=o_first, last=o_last,
                        unpinned=np.array('resize weights + float->u8 rounding restate scikit-image 0.15'))

    # ---- weighted CE: reference class vs oracle -----------------------------------------------------------------
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(2, 3, 24, 40, generator=g) * 2.0
    target = torch.from_numpy(np.stack([synth.class_mask(24, 40, s) for s in (1, 2)])).long()
    w = torch.tensor(olosses.DEFAULT_WEIGHTS)
    assert torch.equal(ref_utils.get_pos_weight(), w)
    p = logits.clone().requires_grad_(True)
    ref_loss = ref_utils.CustomWeightedCrossEntropy(w)(p, target)
    ref_loss.backward()
    o_loss, o_grad = olosses.custom_weighted_cross_entropy_with_grad(logits, target, w)
    assert torch.equal(ref_loss.detach(), o_loss) and torch.equal(p.grad, o_grad)
    report.append('weighted CE: oracle loss and gradient bit-identical to reference utils.CustomWeightedCrossEntropy')
    np.savez_compressed(os.path.join(OUT, 'wce_small.npz'), logits=logits.numpy(), target=target.numpy().astype(np.uint8),
                        weights=w.numpy(), loss=ref_loss.detach().numpy(), grad=p.grad.numpy())

    # ---- dataset enumeration + folder scaffolding ----------------------------------------------------------------
    with tempfile.TemporaryDirectory() as root:
        names = {'sapin': ['b.bmp', 'a.bmp', 'bmp_scan.bmp'], 'epinette_gelee': ['z.bmp', 'c.png'],
                 'other_wood': ['ignored.bmp'], 'epinette_non_gelee': ['m.BMP', 'notes.txt']}
        from PIL import Image
        for wood, fl in names.items():
            os.makedirs(os.path.join(root, 'samples', wood))
            for fn in fl:
                if fn.endswith('.txt'):
                    open(os.path.join(root, 'samples', wood, fn), 'w').write('x')
                else:
                    Image.fromarray(np.zeros((4, 4, 3), np.uint8)).save(os.path.join(root, 'samples', wood, fn),
                                                                        format='BMP' if fn.lower().endswith('bmp') else 'PNG')
        ref_items = ref_dataset.make_dataset_for_dir(root, ref_dataset.IMG_EXTENSIONS)
        ora_items = opipe.make_dataset_for_dir(root)
        ref_norm = [(os.path.relpath(a, root), c, d) for (a, b, c, d) in ref_items]
        ora_norm = [(os.path.relpath(a, root), c, d) for (a, c, d) in ora_items]
        assert ref_norm == ora_norm, (ref_norm, ora_norm)
        ref_predict.generate_folders(root, False)
        ref_dirs = sorted(os.path.relpath(os.path.join(dp, d), root) for dp, dn, _ in os.walk(root) for d in dn)
    with tempfile.TemporaryDirectory() as root2:
        for wood, fl in names.items():
            os.makedirs(os.path.join(root2, 'samples', wood))
        opipe.generate_folders(root2, False)
        ora_dirs = sorted(os.path.relpath(os.path.join(dp, d), root2) for dp, dn, _ in os.walk(root2) for d in dn)
    assert ref_dirs == ora_dirs, (ref_dirs, ora_dirs)
    report.append('dataset order + folder layout: oracle == reference (%d items, %d dirs)' % (len(ref_norm), len(ref_dirs)))
    import json
    json.dump({'names': names, 'items': ref_norm, 'dirs': ref_dirs}, open(os.path.join(OUT, 'dataset_order.json'), 'w'),
              indent=1)

    # ---- stats strings (models.py:323-332): reference arithmetic is inline in _predict_images, restated -------------
    # ---- N1 / N2: Lovasz-Softmax loss (+ gradient), MixedLoss, iou / miou -- the reference's own lovasz_losses module ----
    import lovasz_losses as ref_lovasz
    from oracle import lovasz as olovasz
    g = torch.Generator().manual_seed(8)
    lv_logits = torch.randn(2, 3, 24, 40, generator=g) * 2.0
    lv_target = torch.from_numpy(np.stack([synth.class_mask(24, 40, s) for s in (5, 6)])).long()
    p = lv_logits.clone().requires_grad_(True)
    ref_l = ref_lovasz.LovaszSoftmax()(p, lv_target)
    ref_l.backward()
    ora_l, ora_g = olovasz.lovasz_softmax_with_grad(lv_logits, lv_target)
    assert float(ref_l) == float(ora_l) and torch.equal(p.grad, ora_g), 'Lovasz oracle differs from the reference'
    two = lv_target.clone()
    two[two == 2] = 1                                   # a batch without class 2: classes='present' drops it
    ref_l2 = ref_lovasz.LovaszSoftmax()(lv_logits, two)
    ora_l2, ora_g2 = olovasz.lovasz_softmax_with_grad(lv_logits, two)
    assert float(ref_l2) == float(ora_l2)
    ref_iou = ref_lovasz.iou(lv_logits, lv_target)
    assert np.array_equal(ref_iou, olovasz.iou(lv_logits, lv_target)) and ref_lovasz.miou(lv_logits, lv_target) == olovasz.miou(lv_logits, lv_target)
    w = torch.tensor(olosses.DEFAULT_WEIGHTS)
    ref_mixed = ref_utils.MixedLoss(w)(lv_logits, lv_target) if hasattr(ref_utils, 'MixedLoss') else None
    ora_mixed = olovasz.mixed_loss(lv_logits, lv_target, w)
    if ref_mixed is not None:
        assert float(ref_mixed) == float(ora_mixed), 'MixedLoss oracle differs from the reference'
    np.savez_compressed(os.path.join(OUT, 'lovasz_small.npz'), logits=lv_logits.numpy(), target=lv_target.numpy().astype(np.uint8),
                        loss=np.float32(ora_l), grad=ora_g.numpy(), loss_two_classes=np.float32(ora_l2), grad_two_classes=ora_g2.numpy(),
                        iou=ref_iou, mixed=np.float32(ora_mixed),
                        confusion=olovasz.confusion_matrix(torch.argmax(lv_logits, 1).numpy(), lv_target.numpy()))
    report.append('lovasz: LovaszSoftmax loss %.8f and gradient bit-identical to reference lovasz_losses.LovaszSoftmax; '
                  'iou/miou identical; MixedLoss %s' % (float(ora_l), 'identical to reference utils.MixedLoss' if ref_mixed is not None else 'not importable'))

    m = synth.class_mask(611, 1024, 9)
    np.savez_compressed(os.path.join(OUT, 'stats_small.npz'), mask=m, strings=np.array(opost.class_stats_strings(m)))

    # ---- remove_small_zones: oracle (scipy) vs independent brute force; skimage itself is unavailable ---------------
    rng = np.random.default_rng(0)
    m = synth.class_mask(96, 128, 4).astype(np.int64)
    # blobs of exactly 149 / 150 / 151 px of class 1 and of class 0 inside class 1, plus diagonal (8-conn) links
    m[:40, :] = 0
    for k, n in enumerate((149, 150, 151)):
        blob = np.zeros(160, bool); blob[:n] = True
        m[2:12, 4 + 20 * k: 20 + 20 * k][blob.reshape(10, 16)] = 1 + (k % 2)
    m[50:90, 10:120] = 1
    for k, n in enumerate((149, 150, 151)):
        hole = np.zeros(160, bool); hole[:n] = True
        m[55:65, 14 + 20 * k: 30 + 20 * k][hole.reshape(10, 16)] = 0
    for i in range(30):
        m[20 + (i % 15), 70 + i] = 2 if (i % 2) else 1      # thin diagonal line: 8-connected, < 150 px
    sp = rng.integers(0, 96, 60), rng.integers(0, 128, 60)
    m[sp] = (m[sp] + 1) % 3
    a = opost.remove_small_zones_2d(m)
    b = opost.remove_small_zones_bruteforce(m)
    assert np.array_equal(a, b)
    report.append('remove_small_zones: scipy restatement == brute-force flood fill (unpinned vs scikit-image 0.15)')
    np.savez_compressed(os.path.join(OUT, 'ccl_small.npz'), mask=m.astype(np.uint8), out=a.astype(np.uint8),
                        unpinned=np.array('restates scikit-image 0.15 remove_small_holes/objects'))

    # ---- training loader bookkeeping (N4): the reference's own get_splits + torch's weighted batch sampler ---------------
    rng = np.random.default_rng(3)
    woods = ['epinette_gelee'] * 13 + ['epinette_non_gelee'] * 17 + ['sapin'] * 7
    fake = []
    for i, w in enumerate(woods):
        t = torch.from_numpy((rng.random((24, 32)) < rng.uniform(0.05, 0.6)).astype(np.int64) * rng.integers(1, 3, (24, 32)))
        fake.append((None, t, 'img%d.png' % i, w))
    np.random.seed(7)
    tr, va, te, tw = ref_utils.get_splits(fake)
    from torch.utils.data import BatchSampler, WeightedRandomSampler
    gen = torch.Generator().manual_seed(5)
    batches = list(BatchSampler(WeightedRandomSampler(tw, num_samples=len(tw) * 12, replacement=True, generator=gen),
                                batch_size=5, drop_last=True))                                    # __main__.py:165-168
    np.savez_compressed(os.path.join(OUT, 'splits_small.npz'), woods=np.array(woods), seed=7, sampler_seed=5,
                        label_pixels=np.array([int((t != 0).sum()) for _, t, _, _ in fake]), targets=np.stack([t.numpy() for _, t, _, _ in fake]).astype(np.uint8),
                        train=tr, valid=va, test=te, weights=tw, batches=tr[np.asarray(batches)])
    report.append('get_splits / weighted epoch sampler: fixture written by reference utils.get_splits and torch BatchSampler('
                  'WeightedRandomSampler) (%d / %d / %d items, %d batches of 5)' % (len(tr), len(va), len(te), len(batches)))

    open(os.path.join(OUT, 'PINNING.txt'), 'w').write('\n'.join(report) + '\n')
    print('\n'.join(report))


if __name__ == '__main__':
    main()
