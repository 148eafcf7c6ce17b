"""nanomotif_b200 -- B200-native (sm_100a) implementation of nanomotif's motif-scoring hot path.

Importing the package loads libnmb200.so (built in-tree by ``python nanomotif_b200/build.py``) and
fails loudly when it is missing: there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is absent)
from .api import (  # noqa: F401
    BinContext,
    BinScorer,
    MultiBinScorer,
    clear_caches,
    get_parent_scores,
    methylated_motif_occourances,
    motif_model_bin,
    motif_model_bin_many,
    motif_model_contig,
    subseq_indices,
)
from .device import DeviceAssembly, DevicePileup, MotifPrograms, scan_count  # noqa: F401
from . import dataload, growth, pattern, pipeline, search, sharding, sweep, tables  # noqa: F401
from .model import BetaBernoulliModel, predictive_evaluation_score  # noqa: F401
from .motif import Motif  # noqa: F401
from .pileup import PileupTable  # noqa: F401

__version__ = "0.1.0"
