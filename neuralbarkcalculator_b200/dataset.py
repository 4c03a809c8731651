"""Dataset enumeration + image decode, mirroring the reference's ``dataset.py`` contract.

Kept: the wood-type order, sorted file names, the ``bmp`` -> ``png`` name rewrite, the extension list and the error
behaviour of ``make_dataset_for_dir`` / ``RegressionDatasetFolder`` (reference dataset.py:41-68, 121-149).
Changed: samples are handed over as **uint8 HWC arrays** (no ToTensor float conversion on the host -- the /255 and
Normalize of dataset.py:181-190 + models.py:233-237 happen inside the CUDA stem kernel), and a 24-bit BMP is not
decoded at all: its bottom-up BGR pixel array is passed to the resize kernel as is."""
import os
import random
import struct

import numpy as np
from PIL import Image

WOOD_TYPES = ["epinette_gelee", "epinette_non_gelee", "sapin"]  # dataset.py:50
IMG_EXTENSIONS = ['.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.pgm', '.tif', '.tiff', 'webp']  # dataset.py:76-78


def has_file_allowed_extension(filename, extensions):
    return filename.lower().endswith(tuple(extensions))


def _entries_of(samples_dir, targets_dir, wood_type, extensions):
    """One wood type: every file of every walked sub-folder, joined to the TOP folder as the reference does."""
    top = os.path.join(samples_dir, wood_type)
    for _root, _dirs, names in sorted(os.walk(top)):
        for name in sorted(n for n in names if has_file_allowed_extension(n, extensions)):
            out_name = name.replace("bmp", "png")           # anywhere in the name (dataset.py:58)
            dual = os.path.join(targets_dir, wood_type, out_name)
            yield (os.path.join(top, name), dual if os.path.isfile(dual) else "", out_name, wood_type)


def make_dataset_for_dir(dir, extensions=IMG_EXTENSIONS):
    """(sample_path, target_path or "", fname, wood_type) in the reference's order (dataset.py:41-68):
    wood types in the fixed order above, then sorted file names."""
    samples_dir = os.path.join(dir, "samples")
    if not os.path.isdir(samples_dir):
        raise IOError("Root folder should have a 'samples' subfolder !")
    targets_dir = os.path.join(dir, "duals")
    return [e for wood in WOOD_TYPES for e in _entries_of(samples_dir, targets_dir, wood, extensions)]


def make_dataset(dir, extensions=IMG_EXTENSIONS):
    return make_dataset_for_dir(os.path.expanduser(dir), extensions)


def pil_loader(path, grayscale=False):
    """dataset.py:82-90, returning a uint8 numpy array (HWC RGB or HW)."""
    if not os.path.isfile(path):
        return None
    with open(path, 'rb') as f:
        img = Image.open(f)
        return np.asarray(img.convert('L' if grayscale else 'RGB'))


def read_bmp_pixels(path):
    """Raw pixel array of an uncompressed 24-bit BMP without decoding it.

    Returns (buf uint8 1-D, H, W, pitch, bgr=True, bottom_up) or None when the file is not such a BMP (the caller
    then falls back to PIL, which yields RGB top-down)."""
    with open(path, 'rb') as f:
        head = f.read(54)
        if len(head) < 54 or head[:2] != b'BM':
            return None
        off = struct.unpack_from('<I', head, 10)[0]
        hdr_size, w, h, planes, bpp, comp = struct.unpack_from('<IiiHHI', head, 14)
        if hdr_size < 40 or bpp != 24 or comp != 0 or planes != 1 or w <= 0 or h == 0:
            return None
        bottom_up = h > 0
        h = abs(h)
        pitch = (w * 3 + 3) & ~3
        f.seek(off)
        buf = np.fromfile(f, dtype=np.uint8, count=pitch * h)
        if buf.size != pitch * h:
            return None
    return buf, h, w, pitch, True, bottom_up


class RegressionDatasetFolder:
    """Index-able folder dataset with the reference's constructor arguments (dataset.py:93-149).

    Without transforms ``__getitem__`` hands out what the GPU path consumes: ``(sample_u8_hwc, target_u8_or_None[, fname,
    wood_type])`` -- raw uint8 arrays; /255, Normalize and the label rounding happen in the CUDA kernels
    (``augment.augment_batch`` is the native training loader).  With transforms it behaves as dataset.py:162-205: ONE random
    draw shared by image and label (the generators -- python ``random``, as the reference seeds, and torch's, which current
    torchvision transforms draw from -- are re-seeded with the same number before each of the two ``transform`` calls, so
    a RandomCrop / flip cuts both identically), then ``input_only_transform`` on the image, then, when the transforms
    produced tensors (``ToTensor``), the reference's label conversion: /255 if max > 200, ``round(target * 2)`` as a long
    class map, and a zero map of the image's size when the image has no dual."""

    def __init__(self, root, extensions=IMG_EXTENSIONS, loader=pil_loader, transform=None, input_only_transform=None,
                 include_fname=False, in_memory=False):
        samples = make_dataset(root, extensions)
        if len(samples) == 0:
            raise RuntimeError("Found 0 files in subfolders of: " + root + "\n"
                               "Supported extensions are: " + ",".join(extensions))
        self.root = root
        self.loader = loader
        self.extensions = extensions
        self.transform = transform
        self.input_only_transform = input_only_transform
        self.include_fname = include_fname
        self.in_memory = in_memory
        self.filenames = samples
        self.samples = [(self.loader(p), self.loader(t, grayscale=True) if t else None, f, w)
                        for p, t, f, w in samples] if in_memory else samples

    def __getitem__(self, index):
        sample, target, fname, wood_type = self.samples[index]
        if not self.in_memory:
            sample = self.loader(sample)
            target = self.loader(target, grayscale=True) if target else None
        if self.transform is not None:
            import torch
            random_seed = np.random.randint(2147483647)
            random.seed(random_seed)
            torch.manual_seed(random_seed)
            sample = self.transform(sample)
            if target is not None:
                random.seed(random_seed)
                torch.manual_seed(random_seed)
                target = self.transform(target)
        if self.input_only_transform is not None:
            sample = self.input_only_transform(sample)
        if (self.transform is not None or self.input_only_transform is not None) and hasattr(sample, 'dim'):
            import torch                                   # tensors: the reference's conversions, dataset.py:192-203
            if target is not None:
                target = target if hasattr(target, 'dim') else torch.as_tensor(np.asarray(target))
                target = target.float()
                if target.max() > 200:
                    target = target / 255
                if sample.max() > 200:
                    sample = sample / 255
                target = (target * 2).round_().long().squeeze()
            else:
                target = torch.zeros(sample.shape[1], sample.shape[2])
        if self.include_fname:
            return sample, target, fname, wood_type
        return sample, target

    def __len__(self):
        return len(self.samples)

    def print_filenames(self):
        for idx, filename in enumerate(self.filenames):
            print("{}: {}".format(idx, filename[2]))
