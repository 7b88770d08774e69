"""strainer-gan_b200: B200-native (sm_100a) implementation of Strainer-GAN's straining hot path.

Import as ``strainer_b200`` (the shim at the repository root loads this directory, whose name is
not a valid Python identifier)."""
from .api import *  # noqa: F401,F403
from .api import (D64Scorer, get_scorer, order_stats, percentile_device, quantile_device, compact_indices,  # noqa: F401
                  partition_rows, select_below_percentile, divide_dataset_ensemble, strain_scores, zscore_max,
                  synth_images, ae_errors, mean_plus_k_std, strain_shard, MLPScorer, sort_values, dbscan1d_clean_ratio,
                  ResidentSubset, get_mlp_scorer, gmm_fit_device, dbscan_clean_ratio, U8Images, U8ImageDataset,
                  detect_outliers_fixed, detect_outliers_elbow, detect_outliers_ratio, clear_scorer_caches, D28Scorer,
                  get_d28_scorer, scorer_for, PeerComm)
from .train import TrainableD64, accelerate_discriminator  # noqa: F401
from . import _lib  # noqa: F401

__version__ = "0.1.0"
