"""kwiiyatta_b200: B200-native alignment + spectral-mapping hot path behind kwiiyatta's API.

Public surface mirrors the reference's (kwiiyatta/__init__.py:1-22) for this path:
``align`` / ``align_even`` (kwiiyatta/vocoder/align.py), a ``fastdtw`` module, the converter
back-end class ``B200GMMFeatureConverter`` and ``MLPG``.  See INTEGRATION.md.
"""
from . import fastdtw
from .alignment import (align, align_even, align_even_many, dtw_feature, dtw_feature_many,
                    make_feature, project_path_iter, set_pad_silence)
from .delta import DELTA_WINDOWS, delta_features
from .gmm import B200GMMFeatureConverter, FeatureConverter, GaussianMixture
from .mlpg import MLPG
from .dataset import joint_array_from_pairs, make_dataset_to_array

name = 'kwiiyatta_b200'

__all__ = ['fastdtw', 'align', 'align_even', 'align_even_many', 'dtw_feature',
           'dtw_feature_many', 'make_feature', 'project_path_iter', 'set_pad_silence', 'DELTA_WINDOWS',
           'delta_features', 'B200GMMFeatureConverter', 'FeatureConverter', 'GaussianMixture',
           'MLPG', 'joint_array_from_pairs', 'make_dataset_to_array']
