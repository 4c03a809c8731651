"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol of include/nbc.h,
the Python mirror keeps the reference's interface, and nothing computes without a GPU."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'nbc.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nbc_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_header_symbol(libnbc):
    from neuralbarkcalculator_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(libnbc, n), 'libnbc.so does not export %s' % n
    assert sorted(_lib.SIGNATURES.keys()) == names, 'ctypes table and include/nbc.h disagree'
    assert libnbc.nbc_version() == 100


def test_no_cpu_fallback(libnbc):
    from neuralbarkcalculator_b200 import _lib, ops, utils
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    assert libnbc.nbc_device_check(0) == -3
    assert 'no CUDA device' in _lib.last_error()
    with pytest.raises(RuntimeError):
        utils.remove_small_zones(torch.zeros(1, 8, 8, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        ops.upsample_argmax(torch.zeros(1, 3, 4, 4), (32, 32))
    with pytest.raises(RuntimeError):
        utils.CustomWeightedCrossEntropy(utils.get_pos_weight())(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4).long())


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'neuralbarkcalculator_b200')
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', txt, flags=re.M), fn
                assert '/root/reference' not in txt, fn


def test_model_state_dict_is_drop_in(synthetic_sd):
    import neuralbarkcalculator_b200 as nbc
    m = nbc.fcn_resnet50(pretrained=False)
    assert list(m.state_dict().keys()) == list(synthetic_sd.keys())
    m.load_state_dict(synthetic_sd, strict=True)
    assert hasattr(m, 'backbone') and hasattr(m, 'classifier')
    # torchvision's own FCN-ResNet50 checkpoint layout (num_classes=3, no aux head) is the same key set
    from torchvision.models.segmentation import fcn_resnet50 as tv_fcn
    tv = tv_fcn(weights=None, weights_backbone=None, num_classes=3, aux_loss=False)
    assert set(tv.state_dict().keys()) == set(m.state_dict().keys())
    m.eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 32, 32))       # CPU tensor -> loud failure, not a torch fallback
    m.train()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 32, 32))


def test_dataset_order_matches_reference(golden_dir, tmp_path):
    from PIL import Image
    from neuralbarkcalculator_b200 import dataset, predict
    g = json.load(open(os.path.join(golden_dir, 'dataset_order.json')))
    for wood, files in g['names'].items():
        os.makedirs(tmp_path / 'samples' / wood)
        for fn in files:
            p = tmp_path / 'samples' / wood / fn
            if fn.endswith('.txt'):
                p.write_text('x')
            else:
                Image.fromarray(np.zeros((4, 4, 3), np.uint8)).save(p, format='BMP' if fn.lower().endswith('bmp') else 'PNG')
    items = [(os.path.relpath(a, tmp_path), c, d) for a, b, c, d in dataset.make_dataset_for_dir(str(tmp_path))]
    assert items == [tuple(x) for x in g['items']]
    predict.generate_folders(str(tmp_path), False)
    dirs = sorted(os.path.relpath(os.path.join(dp, d), tmp_path) for dp, dn, _ in os.walk(tmp_path) for d in dn)
    assert dirs == g['dirs']
    with pytest.raises(IOError):
        dataset.make_dataset_for_dir(str(tmp_path / 'nowhere'))
    os.makedirs(tmp_path / 'empty' / 'samples' / 'sapin')
    with pytest.raises(RuntimeError):
        dataset.RegressionDatasetFolder(str(tmp_path / 'empty'))


def test_bmp_raw_reader(tmp_path):
    from oracle import synth
    from neuralbarkcalculator_b200 import dataset
    img = synth.texture_u8(8, 12, 1)
    p = str(tmp_path / 'x.bmp')
    synth.write_bmp(p, img)
    buf, H, W, pitch, bgr, bottom_up = dataset.read_bmp_pixels(p)
    assert (H, W, pitch, bgr, bottom_up) == (8, 12, 36, True, True)
    dec = buf.reshape(H, pitch)[:, :W * 3].reshape(H, W, 3)[::-1, :, ::-1]
    assert np.array_equal(dec, img)


def test_cli_arguments():
    from neuralbarkcalculator_b200 import predict
    a = predict.parse_args(['some/root', '--exclude_nodes'])
    assert a.root_path == 'some/root' and a.exclude_nodes and not a.only_preprocess and a.device == 'cuda:0'
    with pytest.raises(SystemExit):
        predict.parse_args(['some/root', '--device', 'cpu'])


def test_stats_strings_match_reference_arithmetic(golden_dir):
    """The CSV strings computed from integer counts equal the reference's float32 tensor arithmetic."""
    import neuralbarkcalculator_b200 as nbc
    g = np.load(os.path.join(golden_dir, 'stats_small.npz'))
    mask = g['mask']
    counts = np.bincount(mask.ravel(), minlength=3)
    calc = nbc.NeuralBarkCalculator.__new__(nbc.NeuralBarkCalculator)
    calc.mm_per_pix = nbc.NeuralBarkCalculator.DEFAULT_MM_PER_PIXEL
    assert calc._stats_strings(counts.tolist(), mask.size) == list(g['strings'])


def test_png_writer_roundtrip(tmp_path):
    """The pipeline's own PNG encoder (RGB with Sub filter, grey masks, stored / fast / default deflate) decodes to the
    same pixels with PIL."""
    import io
    from PIL import Image
    from neuralbarkcalculator_b200 import _png
    from oracle import synth
    img = synth.texture_u8(37, 53, 1)
    mask = (synth.class_mask(37, 53, 2) * 127).astype(np.uint8)
    for level in (0, 1, 6):
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(_png.encode_png(img, level)))), img)
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(_png.encode_png(mask, level)))), mask)
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(_png.encode_png(img[:, :1], 1)))), img[:, :1])
    _png.write_png(str(tmp_path / 'a.png'), img)
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / 'a.png'))), img)
    with pytest.raises(ValueError):
        _png.encode_png(img.astype(np.float32))


def _png_idat_pixels(data, h, w, ch):
    """Independent decoder of a PNG made by _png.encode_png: zlib inflate + undo filter 0 / 1 (numpy)."""
    import struct
    import zlib
    assert data[:8] == b'\x89PNG\r\n\x1a\n'
    pos, idat = 8, b''
    while pos < len(data):
        n, tag = struct.unpack('>I4s', data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack('>I', data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
        if tag == b'IDAT':
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, w * ch + 1)      # checks the Adler-32 too
    out = raw[:, 1:].copy()
    for r in range(h):
        if raw[r, 0] == 1:
            px = out[r].reshape(w, ch).astype(np.uint32)
            out[r] = (np.cumsum(px, axis=0) & 0xFF).astype(np.uint8).reshape(-1)
        else:
            assert raw[r, 0] == 0
    return out.reshape((h, w, ch) if ch == 3 else (h, w))


def test_native_png_encoder(libnbc):
    """nbc_png_idat (host code in libnbc.so: run-length + per-segment Huffman deflate) against zlib's inflate and against
    the zlib Z_RLE encoder it replaces: same pixels for every edge case, size within 2 %."""
    import ctypes as C
    from neuralbarkcalculator_b200 import _png
    from oracle import synth
    rng = np.random.default_rng(0)
    cases = {
        'texture rgb': synth.texture_u8(208, 333, 3),                              # several segments (21 rows each at w=1024; here 65)
        'texture rgb 1024 wide': synth.texture_u8(64, 1024, 5),
        'class mask': np.array([0, 127, 255], dtype=np.uint8)[synth.class_mask(300, 512, 1)],
        'noise rgb (stored fallback)': rng.integers(0, 256, (50, 400, 3), dtype=np.uint8),
        'noise grey': rng.integers(0, 256, (70, 1000), dtype=np.uint8),
        'zeros': np.zeros((100, 70), np.uint8),
        'constant rgb': np.full((300, 300, 3), 200, np.uint8),                     # runs longer than 258 and across rows
        '1x1 rgb': np.full((1, 1, 3), 7, np.uint8),
        '1x1 grey': np.full((1, 1), 9, np.uint8),
        'one column rgb': rng.integers(0, 4, (500, 1, 3), dtype=np.uint8),
        'one row': rng.integers(0, 3, (1, 5000), dtype=np.uint8),
        'row wider than a stored block': rng.integers(0, 2, (3, 30000, 3), dtype=np.uint8),
        'noise row wider than a stored block': rng.integers(0, 256, (2, 70000), dtype=np.uint8),
        'two symbols': (rng.integers(0, 2, (64, 64), dtype=np.uint8) * 255),
        'skewed histogram (deep Huffman tree)': np.minimum(rng.geometric(0.5, (200, 320)), 255).astype(np.uint8),
    }
    fib = [1, 1]
    while len(fib) < 21:
        fib.append(fib[-1] + fib[-2])
    deep = np.concatenate([np.full(f, k, np.uint8) for k, f in enumerate(fib)])    # Fibonacci counts: a depth-20 Huffman tree
    cases['fibonacci histogram (code lengths must be limited to 15)'] = rng.permutation(deep).reshape(1, -1)
    for name, a in cases.items():
        h, w = a.shape[:2]
        ch = 3 if a.ndim == 3 else 1
        png = _png.encode_png(a, 1)
        assert np.array_equal(_png_idat_pixels(png, h, w, ch), a), name
        ref = _png.encode_png_zlib_rle(a)
        assert np.array_equal(_png_idat_pixels(ref, h, w, ch), a), name
        segments = -(-(h * (w * ch + 1)) // 65535) + h // max(1, 65535 // (w * ch + 1))      # ~160 header bytes each
        assert len(png) <= len(ref) * 1.02 + 200 * (segments + 1), (name, len(png), len(ref))
    # strided input (a window of a larger canvas), and the error paths
    canvas = synth.texture_u8(40, 100, 2)
    win = canvas[:, 10:60]                                                          # row stride 300 bytes, 150 meaningful
    cap = libnbc.nbc_png_idat_bound(40, 50, 3)
    out = np.empty(cap, dtype=np.uint8)
    n = libnbc.nbc_png_idat(C.c_void_p(win.ctypes.data), 40, 50, 3, 300, None, C.c_void_p(out.ctypes.data), cap)
    dense = np.empty(cap, dtype=np.uint8)
    win_c = np.ascontiguousarray(win)
    n2 = libnbc.nbc_png_idat(C.c_void_p(win_c.ctypes.data), 40, 50, 3, 0, None, C.c_void_p(dense.ctypes.data), cap)
    assert n == n2 > 0 and np.array_equal(out[:n], dense[:n2])
    assert libnbc.nbc_png_idat(C.c_void_p(win_c.ctypes.data), 40, 50, 3, 0, None, C.c_void_p(out.ctypes.data), 100) < 0
    assert b'need' in libnbc.nbc_last_error()
    assert libnbc.nbc_png_idat(C.c_void_p(win_c.ctypes.data), 40, 50, 2, 0, None, C.c_void_p(out.ctypes.data), cap) < 0
    assert libnbc.nbc_png_idat_bound(0, 5, 3) == 0
    # lookup table while encoding (dual image 0/127/255 from the class mask, models.py:349-353), file writer, compositor
    from neuralbarkcalculator_b200 import pipeline
    mask = synth.class_mask(123, 77, 4).astype(np.uint8)
    for level in (0, 1, 6):
        png = _png.encode_png(mask, level, lut=pipeline._DUAL_LUT)
        assert np.array_equal(_png_idat_pixels(png, 123, 77, 1), pipeline._DUAL_LUT[mask])
    lut = np.zeros(256, np.uint8)
    assert libnbc.nbc_png_idat(C.c_void_p(win_c.ctypes.data), 40, 50, 3, 0, C.c_void_p(lut.ctypes.data), C.c_void_p(out.ctypes.data), cap) < 0
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        for a, kw in ((cases['texture rgb'], {}), (mask, {'lut': pipeline._DUAL_LUT})):
            _png.write_png(os.path.join(d, 'a.png'), a, 1, **kw)
            assert open(os.path.join(d, 'a.png'), 'rb').read() == _png.encode_png(a, 1, **kw)
    proc = synth.texture_u8(123, 77, 6)
    got = pipeline.combined_image(proc, mask, 'Bark : 12.345;  Node : 0.123   (x.png)')
    assert got.shape == (62 + 16, 2 * 39 + 8, 3)
    assert np.array_equal(got[16:, :39], proc[::2, ::2]) and np.all(got[16:, 39:47] == 255)
    assert np.array_equal(got[16:, 47:], pipeline._VIRIDIS3[mask[::2, ::2]])
    assert (got[:16] == 255).mean() > 0.5 and (got[:16] != 255).any()      # a white strip with the title drawn on it


def test_pipeline_and_numa_host_helpers(tmp_path):
    """Host-side pieces of the folder pipeline and of the one-process-per-GPU plumbing (no GPU needed)."""
    import struct
    from neuralbarkcalculator_b200 import augment, distributed as ndist, pipeline
    from oracle import synth
    # BMP geometry probe: only uncompressed 24-bit 4096x4096 files qualify for the streaming pipeline
    small = str(tmp_path / 's.bmp')
    synth.write_bmp(small, synth.texture_u8(8, 8, 0))
    assert pipeline.bmp_geometry(small) is None and pipeline.bmp_geometry(str(tmp_path / 'missing.bmp')) is None
    big = str(tmp_path / 'b.bmp')
    with open(small, 'rb') as f:
        head = bytearray(f.read(54))
    struct.pack_into('<ii', head, 18, 4096, 4096)          # same header, claimed size 4096 x 4096, bottom-up
    with open(big, 'wb') as f:
        f.write(bytes(head))
    assert pipeline.bmp_geometry(big) == (54, True)
    assert not pipeline.supported([(small, None, 's.png', 'sapin')]) and not pipeline.supported([])
    title_img = pipeline.combined_image(synth.texture_u8(20, 32, 1), synth.class_mask(20, 32, 2), 'Bark : 1.000')
    assert title_img.shape == (10 + 16, 2 * 16 + 8, 3)
    # NUMA helpers degrade gracefully without sysfs / without a GPU
    assert ndist._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert ndist.bind_to_gpu_numa(0)['bound'] in (True, False)
    # augmentation draws respect the transforms' ranges
    p = augment.draw_params(np.random.default_rng(0), 64, 5, 512)
    assert p['x0'].max() <= 512 and p['y0'].min() >= 0 and set(np.unique(p['order'])) <= {0, 1} and p['src'].max() < 5
    assert 0.9 <= p['brightness'].min() and p['brightness'].max() <= 1.1 and 0.8 <= p['saturation'].min() and p['saturation'].max() <= 1.2


def test_profile_tools_read_the_committed_launch_list():
    """tools/kernel_shares.py on the committed ncu launch list: the timed step of the bench is 176 launches, 100 of them
    tensor-core convolutions (49 per chunk of 8 scans + the stem), and the shares add up."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, 'profiles', 'r01u_launches_bench_batch16.csv')
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'kernel_shares.py'), csv_path, '176', '--from', '475'],
                         capture_output=True, text=True, check=True).stdout
    rows = {l.split()[0]: l.split() for l in out.splitlines() if 'launches' in l and not l.startswith('total')}
    assert int(rows['conv_tc_pair_kernel'][1]) + int(rows['conv_tc_kernel'][1]) == 100
    assert abs(sum(float(r[-2]) for r in rows.values()) - 100.0) < 0.5
    assert 'total' in out and '176 launches' in out


def test_integration_doc_calls_match_the_abi():
    """Every `_nbc.nbc_*(...)` call in INTEGRATION.md's binding snippets passes as many arguments as the C-ABI takes
    (`_lib.SIGNATURES`, itself cross-checked against include/nbc.h above) -- a maintainer who copies the snippet must not
    end up passing garbage as a trailing argument."""
    import re
    from neuralbarkcalculator_b200 import _lib
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    calls = 0
    for m in re.finditer(r'_nbc\.(nbc_\w+)\(', text):
        name = m.group(1)
        assert name in _lib.SIGNATURES, 'INTEGRATION.md calls %s, which the ABI does not export' % name
        depth, i, nargs, seen = 1, m.end(), 0, False
        while depth:
            ch = text[i]
            if ch in '([{':
                depth += 1
            elif ch in ')]}':
                depth -= 1
            elif ch == ',' and depth == 1:
                nargs += 1
            elif ch == '#':                      # a comment inside a multi-line call
                i = text.index('\n', i)
                continue
            if depth and not ch.isspace():
                seen = True
            i += 1
        nargs += 1 if seen else 0
        assert nargs == len(_lib.SIGNATURES[name][1]), '%s: INTEGRATION.md passes %d arguments, the ABI takes %d' % (
            name, nargs, len(_lib.SIGNATURES[name][1]))
        calls += 1
    assert calls >= 5


def test_host_zero_row_span():
    """nbc_host_zero_row_span (host code of the library): first / last non-zero memory row, widened to groups of 4."""
    from neuralbarkcalculator_b200 import ops
    H, pitch = 64, 100
    a = np.zeros((H, pitch), np.uint8)
    assert ops.host_zero_row_span(a, H, pitch) == (0, 0)
    a[13, 99] = 1
    assert ops.host_zero_row_span(a, H, pitch) == (12, 4)
    a[13, 99] = 0
    a[13, 97] = 1                                   # beyond row_bytes: padding is not image data
    assert ops.host_zero_row_span(a, H, pitch, row_bytes=96) == (0, 0)
    a[40, 0] = 7
    assert ops.host_zero_row_span(a, H, pitch) == (12, 32)
    a[63, 5] = 1
    a[0, 1] = 1
    assert ops.host_zero_row_span(a, H, pitch) == (0, 64)
    b = np.zeros((10, 8), np.uint8)                 # H not a multiple of the group: the span is clipped to the image
    b[9, 0] = 1
    assert ops.host_zero_row_span(b, 10, 8) == (8, 2)
    t = torch.zeros(32 * 24, dtype=torch.uint8)
    t[5 * 24 + 3] = 1
    assert ops.host_zero_row_span(t, 32, 24) == (4, 4)


def test_dataset_transforms_share_one_random_draw(tmp_path):
    """RegressionDatasetFolder with transforms (dataset.py:162-205): image and label get the SAME random crop / flip, the
    label becomes round(2 * gray / 255) as a long class map, and an image without a dual gets a zero map."""
    from PIL import Image
    import torchvision.transforms as T
    from neuralbarkcalculator_b200.dataset import RegressionDatasetFolder
    root = tmp_path / 'ds'
    for sub in ('samples/sapin', 'duals/sapin'):
        os.makedirs(root / sub)
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 3, (40, 56)).astype(np.uint8)
    img = np.stack([lab * 100, lab * 100 + 1, lab * 100 + 2], axis=-1).astype(np.uint8)      # the image encodes its label
    Image.fromarray(img).save(root / 'samples/sapin/a.png')
    Image.fromarray((lab * 127 + (lab == 2)).astype(np.uint8)).save(root / 'duals/sapin/a.png')   # 0 / 127 / 255
    Image.fromarray(img).save(root / 'samples/sapin/b.png')                                   # no dual
    tf = T.Compose([T.ToPILImage(), T.RandomCrop(24), T.RandomHorizontalFlip(), T.RandomVerticalFlip(), T.ToTensor()])
    ds = RegressionDatasetFolder(str(root), transform=tf, include_fname=True)
    for _ in range(8):
        x, y, fname, wood = ds[0]
        assert x.shape == (3, 24, 24) and y.shape == (24, 24) and y.dtype == torch.int64 and (fname, wood) == ('a.png', 'sapin')
        assert torch.equal((x[0] * 255).round().long() // 100, y), 'image and label were cut differently'
    x, y, _, _ = ds[1]
    assert y.shape == (24, 24) and not y.any()
    raw = RegressionDatasetFolder(str(root))[0]
    assert raw[0].dtype == np.uint8 and raw[0].shape == (40, 56, 3) and raw[1].shape == (40, 56)


def test_get_splits_and_epoch_sampler_match_reference(golden_dir):
    """N4 bookkeeping: get_splits (utils.py:76-132) and the weighted epoch batches (__main__.py:165-172) against a fixture
    written by the reference's own get_splits and torch's BatchSampler(WeightedRandomSampler) (oracle/make_golden.py)."""
    from neuralbarkcalculator_b200 import utils as nu
    g = np.load(os.path.join(golden_dir, 'splits_small.npz'))
    ds = [(None, torch.from_numpy(t.astype(np.int64)), 'img%d.png' % i, str(w)) for i, (t, w) in enumerate(zip(g['targets'], g['woods']))]
    np.random.seed(int(g['seed']))
    tr, va, te, tw = nu.get_splits(ds)
    assert np.array_equal(tr, g['train']) and np.array_equal(va, g['valid']) and np.array_equal(te, g['test'])
    assert tw.dtype == np.float32 and np.array_equal(tw, g['weights'])                       # bit for bit
    np.random.seed(int(g['seed']))
    tr2, _, _, tw2 = nu.get_splits(ds, label_pixels=g['label_pixels'])                      # counts from the GPU loader
    assert np.array_equal(tr2, tr) and np.array_equal(tw2, tw)
    b = nu.weighted_epoch_batches(tr, tw, 5, generator=torch.Generator().manual_seed(int(g['sampler_seed'])))
    assert np.array_equal(b.numpy(), g['batches'])


def test_combined_image_modes(tmp_path, monkeypatch):
    """results/combined_images: NBC_COMBINED selects the native stand-in (default), the reference's matplotlib figure
    ('figure': an optional dependency -- loud error when it is missing, never a silent skip) or nothing."""
    from PIL import Image
    from neuralbarkcalculator_b200 import figure, pipeline
    rng = np.random.default_rng(0)
    proc = rng.integers(0, 255, (40, 64, 3), dtype=np.uint8)
    mask = rng.integers(0, 3, (40, 64), dtype=np.uint8)
    stats = ['12.50000', '1.00000', '0.25000', '2.00000']
    monkeypatch.delenv('NBC_COMBINED', raising=False)
    assert figure.mode() == 'standin'
    out = str(tmp_path / 'c.png')
    pipeline.write_combined(out, proc, mask, stats, 'a.png', [0.7, 0.6, 0.4], [0.1, 0.1, 0.1])
    img = np.asarray(Image.open(out))
    assert img.shape == (20 + 16, 2 * 32 + 8, 3) and np.array_equal(img[16:, :32], proc[::2, ::2])
    monkeypatch.setenv('NBC_COMBINED', '0')
    assert figure.mode() == 'off'
    pipeline.write_combined(str(tmp_path / 'none.png'), proc, mask, stats, 'a.png', [0.7] * 3, [0.1] * 3)
    assert not os.path.exists(str(tmp_path / 'none.png'))
    monkeypatch.setenv('NBC_COMBINED', 'figure')
    assert figure.mode() == 'figure'
    assert figure.suptitle_text((12.5, 0.25)) == 'Estimated composition percentages\nBark : 12.500\nNode : 0.250\n'   # models.py:334-340
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match='matplotlib'):
            pipeline.write_combined(str(tmp_path / 'f.png'), proc, mask, stats, 'a.png', [0.7] * 3, [0.1] * 3)
    else:
        pipeline.write_combined(str(tmp_path / 'f.png'), proc, mask, stats, 'a.png', [0.7] * 3, [0.1] * 3)
        assert os.path.getsize(str(tmp_path / 'f.png')) > 10000


def test_read_scan_copies_only_the_rows_between_the_dark_bands(tmp_path):
    """pipeline.read_scan (host code): a 4096^2 BMP is mapped, its all-zero bands are found in the page cache and only the
    span in between lands in the (pinned) buffer; the span equals nbc_host_zero_row_span of the whole array, and the
    plain-read path gives the same span with the whole array copied."""
    from neuralbarkcalculator_b200 import pipeline
    from oracle import synth
    RAW = pipeline.RAW
    rng = np.random.default_rng(0)
    img = np.zeros((RAW, RAW, 3), np.uint8)
    img[1201:2999] = rng.integers(1, 255, (1798, RAW, 3), dtype=np.uint8)     # band edges that are not multiples of 4
    path = str(tmp_path / 'scan.bmp')
    synth.write_bmp(path, img)
    off, bottom_up = pipeline.bmp_geometry(path)
    pitch = RAW * 3
    buf = torch.full((RAW * RAW * 3,), 0x5A, dtype=torch.uint8)
    row0, rows = pipeline.read_scan(path, off, buf, True)
    whole = torch.empty(RAW * RAW * 3, dtype=torch.uint8)
    assert pipeline.read_scan(path, off, whole, False) == (0, RAW)
    from neuralbarkcalculator_b200 import ops
    assert (row0, rows) == ops.host_zero_row_span(whole, RAW, pitch)
    assert row0 % 4 == 0 and rows % 4 == 0 and 0 < rows < RAW
    assert torch.equal(buf[row0 * pitch:(row0 + rows) * pitch], whole[row0 * pitch:(row0 + rows) * pitch])
    assert bool((buf[:row0 * pitch] == 0x5A).all()) and bool((buf[(row0 + rows) * pitch:] == 0x5A).all())      # never written
    assert not whole[:row0 * pitch].any() and not whole[(row0 + rows) * pitch:].any()                          # ... and all zero
