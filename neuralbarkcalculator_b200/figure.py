"""The reference's "combined image" (models.py:280-347), drawn with matplotlib when it is installed.

The reference saves, for every processed image, a two-panel matplotlib figure under ``results/combined_images/<wood>/``:
the (normalised) network input next to the class map (``imshow(vmax=2)``, i.e. viridis at 0 / 0.5 / 1), a legend with
one patch per class present, the estimated composition as suptitle, ``tight_layout`` and ``savefig(dpi=900)`` --
5760 x 4320 RGBA pixels, about a second of host time per image.  matplotlib is an OPTIONAL dependency of this package:

* ``NBC_COMBINED=figure`` renders exactly that figure here (same calls, same inputs: the f32-normalised image and the
  integer class map), and fails loudly when matplotlib is missing;
* the default (``NBC_COMBINED=1``) writes the native stand-in of ``pipeline.combined_image`` -- same panels, colours and
  numbers at half resolution, encoded by the library's own PNG encoder at hundreds of images per second;
* ``NBC_COMBINED=0`` writes nothing there.
"""
import os

import numpy as np

CLASS_NAMES = ['Nothing', 'Bark', 'Node']        # models.py:286


def mode():
    """'figure' | 'standin' | 'off' from NBC_COMBINED."""
    v = os.environ.get('NBC_COMBINED', '1').lower()
    if v in ('0', 'off', 'no', 'false'):
        return 'off'
    return 'figure' if v in ('figure', 'matplotlib', 'mpl') else 'standin'


def suptitle_text(percents):
    """models.py:334-340: the composition block above the panels; percents = (bark %, node %) as floats."""
    text = 'Estimated composition percentages\n'
    for name, pct in zip(CLASS_NAMES[1:], percents):
        text += '{} : {:.3f}\n'.format(name, pct)
    return text


def save_reference_figure(path, proc_u8, mask, percents, mean, std, dpi=900):
    """Render and save the reference figure.  proc_u8: processed image u8 [H,W,3]; mask: class map [H,W] (after region
    removal / exclude-nodes); percents: (bark %, node %); mean / std: the Normalize constants of models.py:208-209 (the
    reference plots the NORMALISED tensor, which imshow clips to [0, 1])."""
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.patches as mpatches
        import matplotlib.pyplot as plt
    except ImportError as e:
        raise RuntimeError('NBC_COMBINED=figure needs matplotlib (an optional dependency); unset it for the native '
                           'stand-in or set NBC_COMBINED=0') from e
    x = (np.asarray(proc_u8, dtype=np.float32) / np.float32(255) - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
    mask = np.asarray(mask).astype(np.int64)
    fig, axs = plt.subplots(1, 2)
    patches = []
    for ax, img, name in zip(axs.flatten(), (x, mask), ('Input', 'Generated image')):
        shown = ax.imshow(img, vmax=2)
        ax.set_title(name)
        ax.axis('off')
        if img.ndim == 2:      # the predicted classes: one legend entry per class present (models.py:304-311)
            patches = [mpatches.Patch(color=shown.cmap(shown.norm(v)), label='{} zone'.format(CLASS_NAMES[v]))
                       for v in np.unique(img.ravel())]
    fig.legend(handles=patches, title='Classes', bbox_to_anchor=(0.4, -0.2, 0.5, 0.5))
    plt.suptitle(suptitle_text(percents))
    plt.tight_layout()
    plt.savefig(path, format='png', dpi=dpi)
    plt.close()
