"""Oracle for everything after the argmax (reference ``utils.py:135-148``, ``models.py:273-276, 323-364``).

Test infrastructure only.  ``remove_small_zones`` depends on scikit-image 0.15 ``remove_small_holes`` /
``remove_small_objects`` (not installable here -> **parity unpinned**); they are restated from their published
definition: ``ndi.label(ar, generate_binary_structure(ar.ndim, connectivity))`` + ``np.bincount`` +
``component_sizes < min_size`` -> cleared.  ``remove_small_holes(ar, t)`` is ``~remove_small_objects(~ar, t)``.
The build labels each image on its own with the full 3x3 structure (the reference's 3-D structure
only differs for batch > 1, which predict never uses -- models.py:249-250).
"""
import numpy as np
from scipy import ndimage as ndi

STRUCT8 = np.ones((3, 3), dtype=bool)
MM_PER_PIX = 3.6 * 3.6  # models.py:210


def _remove_small_objects(ar, min_size):
    """skimage.morphology.remove_small_objects(ar: bool, min_size, connectivity=2), 2-D, returns a copy."""
    out = ar.copy()
    labels, _ = ndi.label(ar, structure=STRUCT8)
    sizes = np.bincount(labels.ravel())
    too_small = sizes < min_size
    too_small[0] = False  # label 0 is the background of this pass, never "removed"
    out[too_small[labels]] = False
    return out


def remove_small_zones_2d(mask, threshold=150):
    """utils.py:135-148 on one [H, W] integer class mask; returns a new array of the same dtype.

    np_image = (img == 0)
    remove_small_holes(np_image, 150, connectivity=2)     # small non-zero blobs become 'nothing'
    remove_small_objects(np_image, 150, connectivity=2)   # small 'nothing' islands become foreground
    img[(np_image == 0) & (img == 0)] = 1 ; img[(np_image != 0) & (img != 0)] = 0"""
    img = np.array(mask, copy=True)
    bg = (img == 0)
    bg = ~_remove_small_objects(~bg, threshold)   # stage A (remove_small_holes)
    bg = _remove_small_objects(bg, threshold)     # stage B, on the output of A
    img[(~bg) & (img == 0)] = 1
    img[bg & (img != 0)] = 0
    return img


def remove_small_zones(mask, threshold=150):
    """Batched [B, H, W] (or [H, W]) wrapper, per-image labelling."""
    m = np.asarray(mask)
    if m.ndim == 2:
        return remove_small_zones_2d(m, threshold)
    return np.stack([remove_small_zones_2d(x, threshold) for x in m])


def remove_small_zones_bruteforce(mask, threshold=150):
    """Independent pure-Python flood-fill version (no scipy) used to pin the scipy restatement on small cases."""
    img = np.array(mask, copy=True)
    H, W = img.shape

    def components(binary):
        seen = np.zeros_like(binary, dtype=bool)
        comps = []
        for y in range(H):
            for x in range(W):
                if binary[y, x] and not seen[y, x]:
                    stack, comp = [(y, x)], []
                    seen[y, x] = True
                    while stack:
                        cy, cx = stack.pop()
                        comp.append((cy, cx))
                        for dy in (-1, 0, 1):
                            for dx in (-1, 0, 1):
                                ny, nx = cy + dy, cx + dx
                                if 0 <= ny < H and 0 <= nx < W and binary[ny, nx] and not seen[ny, nx]:
                                    seen[ny, nx] = True
                                    stack.append((ny, nx))
                    comps.append(comp)
        return comps

    bg = (img == 0)
    for comp in components(~bg):
        if len(comp) < threshold:
            for (y, x) in comp:
                bg[y, x] = True
    for comp in components(bg.copy()):
        if len(comp) < threshold:
            for (y, x) in comp:
                bg[y, x] = False
    img[(~bg) & (img == 0)] = 1
    img[bg & (img != 0)] = 0
    return img


def exclude_nodes(mask):
    """models.py:273-276: node pixels (2) are rewritten to class 1 ("nothing_class" is 1 in the reference)."""
    out = np.array(mask, copy=True)
    out[out == 2] = 1
    return out


def class_stats_strings(mask, mm_per_pix=MM_PER_PIX):
    """models.py:323-332: for c in (1, 2): '{:.5f}' of mean((mask==c).float())*100 and of sum*mm_per_pix.

    The reference does this with float32 torch tensors: mean -> f32, *100 -> f32 tensor formatted as a float;
    sum -> f32, * python float -> f32 tensor, .item()."""
    import torch
    t = torch.as_tensor(np.asarray(mask))
    out = []
    for c in (1, 2):
        n = (t == c).float()
        pct = n.mean()
        out.append('{:.5f}'.format(pct * 100))
        out.append('{:.5f}'.format((n.sum() * mm_per_pix).item()))
    return out


def dual_image(mask):
    """models.py:349-353: u8 image 0 / 127 / 255."""
    m = np.asarray(mask)
    out = np.zeros(m.shape, dtype=np.uint8)
    out[m == 1] = 127
    out[m == 2] = 255
    return out


CSV_HEADER = ['Name', 'Type', 'Image Size', 'Output Bark %', 'Bark area (mm^2)', 'Output Node %',
              'Node area (mm^2)']  # models.py:252-255 (7 names; rows carry 6 fields, models.py:321)
