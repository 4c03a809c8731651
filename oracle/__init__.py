"""CPU oracle for the NeuralBarkCalculator segmentation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product package
(``neuralbarkcalculator_b200``) never imports it and has no CPU fallback.

Every function restates one piece of the reference (``/root/reference/src/bark_calculator``) and cites
the file:line it follows.  Pinning status (see DESIGN.md "Oracle"):

* model forward / bicubic upsample / argmax / weighted CE / trim_black / dataset enumeration:
  pinned against the reference's own code imported with stubbed plotting / skimage modules
  (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
* 4x cubic resize, float->u8 on save, remove_small_zones: the arithmetic lives in scikit-image 0.15
  (requirements.txt:5), which is not installed here and cannot be fetched -> **parity unpinned** for
  those three pieces; they restate the published algorithm (see each docstring).
"""
