"""Oracle for the preprocessing stage (reference ``models.py:157-203``).  Test infrastructure only.

Reference flow (``Preprocessor._preprocess_image``, models.py:191-203):

    image = CHW float32 (ToTensor: u8/255)  -> HWC
    if max(image.shape) > 1024: image = skimage.transform.resize(image, (1024, 1024), order=3,
                                                                 mode='reflect', anti_aliasing=False)
    if image.shape[0] == image.shape[1]: image = trim_black(image)
    skimage.io.imsave(path, image)          # float -> uint8 PNG

scikit-image 0.15 (requirements.txt:5) is not installable here, so the resize and the float->u8 save are a
restatement of its published algorithm -- **parity unpinned** for these two:

* ``resize`` -> ``warp`` with the pixel-centre-preserving affine map ``src = (dst + 0.5) * scale - 0.5``,
  order-3 = Catmull-Rom cubic convolution (a = -0.5) evaluated per channel in float64, borders by
  ``mode='reflect'`` (numpy "symmetric"), then ``_clip_warp_output``: clip to the input's global [min, max].
  At the exact 4x ratio (4096 -> 1024) ``src = 4 i + 1.5``: the four taps are rows ``4i .. 4i+3`` with weights
  ``[-1, 9, 9, -1] / 16`` -- never touching the border -- so ``256 * out`` is an exact integer combination of
  the u8 input (``S`` below) and the whole stage is integer arithmetic.
* float -> u8 on save: ``round_half_up(255 * x)`` (img_as_ubyte); with ``x = clip(S) / (256 * 255)`` this is
  ``(clip(S) + 128) >> 8``.  Ties (S = 128 mod 256) may differ by one LSB from the reference, whose float32
  ``u8 / 255`` noise decides them.

``trim_black`` is pinned: ``oracle/make_golden.py`` runs the reference's own function on the same inputs.
"""
import numpy as np

TARGET = 1024
W4 = np.array([-1, 9, 9, -1], dtype=np.int64)  # 16 * Catmull-Rom weights at t = 0.5


def resize4x_S(img_u8):
    """Exact integer ``S = 256 * (unclipped cubic resize)`` for a 4x reduction.

    img_u8: [4h, 4w, C] uint8 -> int64 [h, w, C].  Follows models.py:194-198 at scale 4."""
    H, W, C = img_u8.shape
    assert H % 4 == 0 and W % 4 == 0
    x = img_u8.astype(np.int64).reshape(H // 4, 4, W // 4, 4, C)
    return np.einsum('ipjqc,p,q->ijc', x, W4, W4)


def resize_general_f64(img_u8, out_h=TARGET, out_w=TARGET):
    """General-ratio restatement (float64, separable Catmull-Rom, symmetric border, clip to [min,max]).

    Returns float64 in u8 units (0..255).  Evaluated rows first then columns with the
    ``f1 + 0.5 x (f2 - f0 + x (2 f0 - 5 f1 + 4 f2 - f3 + x (3 (f1 - f2) + f3 - f0)))`` form."""
    img = img_u8.astype(np.float64)
    H, W, _ = img.shape

    def taps(n_in, n_out):
        scale = n_in / n_out
        src = (np.arange(n_out) + 0.5) * scale - 0.5
        i0 = np.floor(src).astype(np.int64)
        t = src - i0
        idx = i0[:, None] + np.arange(-1, 3)[None, :]
        # numpy 'symmetric' reflection: d c b a | a b c d | d c b a
        period = 2 * n_in
        idx = np.mod(idx, period)
        idx = np.where(idx >= n_in, period - 1 - idx, idx)
        return idx, t

    def cubic(f0, f1, f2, f3, x):
        return f1 + 0.5 * x * (f2 - f0 + x * (2.0 * f0 - 5.0 * f1 + 4.0 * f2 - f3 + x * (3.0 * (f1 - f2) + f3 - f0)))

    ci, ct = taps(W, out_w)
    ri, rt = taps(H, out_h)
    ct = ct[None, :, None]
    cols = cubic(img[:, ci[:, 0]], img[:, ci[:, 1]], img[:, ci[:, 2]], img[:, ci[:, 3]], ct)
    rt = rt[:, None, None]
    out = cubic(cols[ri[:, 0]], cols[ri[:, 1]], cols[ri[:, 2]], cols[ri[:, 3]], rt)
    return np.clip(out, img.min(), img.max())


def trim_rows_from_counts(nondark_count, width, height):
    """``trim_black`` row rule (models.py:157-166) on per-row counts of non-dark pixels.

    keep[r] = mean(nondark[r, :]) > 0.85; first = argmax(keep); last = H - argmax(keep[::-1])."""
    keep = (nondark_count.astype(np.float64) / width) > 0.85
    first = int(np.argmax(keep))
    last = int(height - np.argmax(keep[::-1]))
    return first, last


def trim_black_float(image):
    """Verbatim restatement of ``trim_black`` (models.py:157-166) on a float HWC image."""
    summed = np.sum(image, axis=-1) > 1e-3
    keep = np.mean(summed, axis=-1) > 0.85
    first = np.argmax(keep)
    last = image.shape[0] - np.argmax(keep[::-1])
    return image[first:last], int(first), int(last)


def preprocess_u8(img_u8, target=TARGET):
    """Full preprocessing of one raw image -> (processed u8 [H', W', 3], first, last).

    Integer restatement of models.py:191-203 (+ skimage save) valid when either the image needs no resize
    or it is an exact 4x reduction to ``target`` (the 4096^2 BMP case of BASELINE.json)."""
    H, W, C = img_u8.shape
    if max(H, W, C) > target:
        if H == 4 * target and W == 4 * target:
            S = resize4x_S(img_u8)
            lo, hi = int(img_u8.min()), int(img_u8.max())
            Sc = np.clip(S, 256 * lo, 256 * hi)
            out = ((Sc + 128) >> 8).astype(np.uint8)
            # float image value = Sc / (256 * 255); "sum over channels > 1e-3"  <=>  sum(Sc) >= 66
            nondark = Sc.sum(axis=-1) * (1.0 / (256.0 * 255.0)) > 1e-3
        else:
            f = resize_general_f64(img_u8, target, target)
            out = np.floor(f + 0.5).astype(np.uint8)
            nondark = (f / 255.0).sum(axis=-1) > 1e-3
    else:
        out = img_u8
        nondark = (img_u8.astype(np.float64) / 255.0).sum(axis=-1) > 1e-3
    h, w = out.shape[:2]
    first, last = 0, h
    if h == w:  # "Untrimmed" (models.py:200)
        first, last = trim_rows_from_counts(nondark.sum(axis=1), w, h)
    return np.ascontiguousarray(out[first:last]), first, last
