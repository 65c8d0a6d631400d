"""Exception types of the hot path (mirror of lib/epipolar/eight_point.py:20-23)."""

EIGHT_POINT_ERROR_MESSAGE = (
    "More than one eigenvalue of Y.T @ Y is small. Cannot confidently estimate"
    " fundamental matrix."
)


class EightPointCalculationError(Exception):
    """Raised if the computation cannot proceed due to ill-conditioned input data."""
