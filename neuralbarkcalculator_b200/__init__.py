"""B200-native (sm_100a) segmentation hot path of NeuralBarkCalculator.

Host side mirrors the reference's Python surface (``models.py``, ``utils.py``, ``dataset.py``, ``predict.py`` of
TortillasAlfred/NeuralBarkCalculator) and calls hand-written CUDA through the C-ABI of ``libnbc.so``
(``include/nbc.h``).  There is no CPU fallback: importing works anywhere, computing needs a B200."""
from .models import (FCNHead, NeuralBarkCalculator, Preprocessor, SimpleSegmentationModel, fcn_resnet50,  # noqa: F401
                     trim_black)
from .utils import CustomWeightedCrossEntropy, get_pos_weight, remove_small_zones  # noqa: F401

__all__ = ['fcn_resnet50', 'SimpleSegmentationModel', 'FCNHead', 'Preprocessor', 'NeuralBarkCalculator', 'trim_black',
           'remove_small_zones', 'CustomWeightedCrossEntropy', 'get_pos_weight']
