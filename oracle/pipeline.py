"""End-to-end CPU restatement of ``predict.py`` (reference ``predict.py:10-58``, ``dataset.py:41-68, 82-90,
162-205``, ``models.py:173-203, 230-364``).  Test infrastructure / timed CPU baseline only.

Not restated: the matplotlib 900-dpi "combined image" (models.py:280-347) -- matplotlib is not installed; leaving
it out flatters the reference's CPU time."""
import csv
import os
import time
import numpy as np
import torch
from PIL import Image

from . import model as omodel
from . import postprocess as opost
from . import preprocess as opre

WOOD_TYPES = ['epinette_gelee', 'epinette_non_gelee', 'sapin']
IMG_EXTENSIONS = ['.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.pgm', '.tif', '.tiff', 'webp']


def make_dataset_for_dir(root):
    """dataset.py:41-68: fixed wood-type order, then sorted file names; 'bmp' -> 'png' anywhere in the name."""
    samples_dir = os.path.join(root, 'samples')
    if not os.path.isdir(samples_dir):
        raise IOError("Root folder should have a 'samples' subfolder !")
    items = []
    for wood in WOOD_TYPES:
        d = os.path.join(samples_dir, wood)
        for _, _, fnames in sorted(os.walk(d)):
            for fname in sorted(fnames):
                if any(fname.lower().endswith(e) for e in IMG_EXTENSIONS):
                    items.append((os.path.join(d, fname), fname.replace('bmp', 'png'), wood))
    return items


def generate_folders(root, only_preprocess=False):
    """predict.py:10-48."""
    present = set(os.listdir(os.path.join(root, 'samples'))) & set(WOOD_TYPES)
    for w in present:
        os.makedirs(os.path.join(root, 'processed', 'samples', w), exist_ok=True)
        if not only_preprocess:
            for lvl in ('combined_images', 'outputs'):
                os.makedirs(os.path.join(root, 'results', lvl, w), exist_ok=True)


def load_rgb(path):
    """dataset.py:82-90."""
    with open(path, 'rb') as f:
        return np.asarray(Image.open(f).convert('RGB'))


def preprocess_folder(root, timers=None):
    items = make_dataset_for_dir(root)
    if not items:
        raise RuntimeError('Found 0 files in subfolders of: ' + root)
    for path, fname, wood in items:
        t0 = time.perf_counter()
        raw = load_rgb(path)
        t1 = time.perf_counter()
        out, _, _ = opre.preprocess_u8(raw)
        t2 = time.perf_counter()
        Image.fromarray(out).save(os.path.join(root, 'processed', 'samples', wood, fname.replace('.bmp', '.png')))
        t3 = time.perf_counter()
        if timers is not None:
            timers['decode'] = timers.get('decode', 0) + t1 - t0
            timers['resize_trim'] = timers.get('resize_trim', 0) + t2 - t1
            timers['png_save'] = timers.get('png_save', 0) + t3 - t2


@torch.no_grad()
def predict_folder(root, model, exclude_nodes=False, timers=None):
    """models.py:230-364 without the matplotlib figure.  Returns the CSV rows (header first)."""
    proc = os.path.join(root, 'processed')
    rows = [list(opost.CSV_HEADER)]
    for path, fname, wood in make_dataset_for_dir(proc):
        t0 = time.perf_counter()
        img = load_rgb(path)
        x = omodel.normalise_u8(img)
        t1 = time.perf_counter()
        logits = model(x)
        t2 = time.perf_counter()
        mask = torch.argmax(logits, dim=1)[0].numpy()
        mask = opost.remove_small_zones_2d(mask)
        if exclude_nodes:
            mask = opost.exclude_nodes(mask)
        t3 = time.perf_counter()
        rows.append([fname, wood] + opost.class_stats_strings(mask))
        Image.fromarray(opost.dual_image(mask), mode='L').save(os.path.join(root, 'results', 'outputs', wood, fname))
        t4 = time.perf_counter()
        if timers is not None:
            for k, v in (('load_norm', t1 - t0), ('forward', t2 - t1), ('argmax_ccl', t3 - t2), ('stats_png', t4 - t3)):
                timers[k] = timers.get(k, 0) + v
    with open(os.path.join(root, 'results', 'final_stats.csv'), 'w') as f:  # models.py:360-364 (no newline='')
        csv.writer(f, delimiter='\t').writerows(rows)
    return rows


def predict_main(root, state_dict, exclude_nodes=False, only_preprocess=False, timers=None):
    """predict.py:51-58."""
    generate_folders(root, only_preprocess)
    preprocess_folder(root, timers)
    if only_preprocess:
        return None
    return predict_folder(root, omodel.load_model(state_dict), exclude_nodes, timers)
