"""B200-native implementation of UniBwTAS/continuous_clustering's per-column hot path (range-image insertion,
ground-point segmentation, point association / cluster-tree merge, finished-cluster detection) behind the
reference's ContinuousClustering class API. CUDA (sm_100a) only -- see DESIGN.md."""
from .api import (  # noqa: F401
    BatchResult,
    ClusteringError,
    Configuration,
    ContinuousClustering,
    ContinuousClusteringConfiguration,
    ContinuousGroundSegmentationConfiguration,
    ContinuousRangeImageConfiguration,
    GeneralConfiguration,
    KittiEvaluation,
    KittiReplay,
    OusterInput,
)
