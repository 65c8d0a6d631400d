"""Mirror of lib/common/feature.py:4-7 (boundary value type)."""
import dataclasses


@dataclasses.dataclass
class Feature:
    x: float
    y: float
