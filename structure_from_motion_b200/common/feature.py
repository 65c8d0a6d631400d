"""The image-feature value type that crosses the hot-path boundary (lib/common/feature.py:4-7): a position in
pixel (or K-normalised) coordinates.  Field order, equality and repr are what callers of the reference rely on."""
from dataclasses import dataclass


@dataclass
class Feature:
    x: float  # column
    y: float  # row
