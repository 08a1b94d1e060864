"""Per-workload configurations: the reference's struct defaults (hpp:24-87) + the ego box of
kitti_demo.cpp:286-291 + the launch-file overrides of the named sensor."""
from __future__ import annotations

from . import synth
from .api import Configuration


def stream_configuration(spec_name: str, **overrides) -> Configuration:
    sp = synth.spec(spec_name)
    cfg = Configuration()
    cfg.general.is_single_threaded = True
    cfg.range_image.num_columns = sp.num_columns
    g = cfg.ground_segmentation
    g.height_ref_to_maximum_ = 0.5
    g.height_ref_to_ground_ = -sp.sensor_height
    g.length_ref_to_front_end_, g.length_ref_to_rear_end_ = 3.0, -3.0
    g.width_ref_to_left_mirror_, g.width_ref_to_right_mirror_ = 1.5, -1.5
    if spec_name == "kitti64":  # kitti_demo.cpp:279-284
        cfg.clustering.ignore_points_in_chessboard_pattern = False
        cfg.clustering.max_distance = 0.5
    if spec_name.startswith("os32"):  # sensor_os32_left.launch:18-27
        g.fog_filtering_intensity_below = 3
        g.fog_filtering_distance_below = 5.0
        g.fog_filtering_inclination_above = -0.17
        cfg.clustering.ignore_points_in_chessboard_pattern = False
        cfg.clustering.ignore_points_with_too_big_inclination_angle_diff = False
    for k, v in overrides.items():
        for group in (cfg.general, cfg.range_image, cfg.ground_segmentation, cfg.clustering):
            if hasattr(group, k):
                setattr(group, k, v)
                break
        else:
            raise AttributeError(k)
    return cfg
