"""Drop-in ``lib`` package: the reference's import paths, served by structure_from_motion_b200.

``from lib.epipolar.epipolar_ransac import estimate_essential_mat_with_ransac`` (apps/sfm.py:14-22
of the reference) resolves to the B200-native implementation; every ``lib.X.Y`` module object
IS the corresponding ``structure_from_motion_b200.X.Y`` module.
"""
import importlib
import sys

_IMPL = "structure_from_motion_b200"
_MODULES = [
    "common", "common.feature", "common.correlate",
    "blur", "blur.gaussian",
    "harris", "harris.harris_detector",
    "feature_matching", "feature_matching.util", "feature_matching.ncc", "feature_matching.ssd",
    "feature_matching.matching",
    "transforms", "transforms.transforms",
    "data_utils", "data_utils.middlebury_utils",
    "ransac", "ransac.ransac",
    "epipolar", "epipolar.triangulation", "epipolar.sed", "epipolar.eight_point",
    "epipolar.epipolar_ransac",
]
for _name in _MODULES:
    _mod = importlib.import_module(f"{_IMPL}.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    _parent, _, _leaf = _name.rpartition(".")
    if _parent:
        setattr(sys.modules[f"{__name__}.{_parent}"], _leaf, _mod)
    else:
        globals()[_leaf] = _mod
