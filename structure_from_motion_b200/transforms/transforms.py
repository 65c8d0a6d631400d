"""Rigid transform value type crossing the hot-path boundary: the API of lib/transforms/transforms.py:10-66
(``Transform3D.from_rmat_t``, ``identity``, ``Tmat`` / ``Rmat`` / ``t``, ``inv``, ``*`` and ``@``), re-implemented.

The 4x4 matrix is assembled directly from R and t (the reference goes through transforms3d.affines.compose with unit
zooms, transforms.py:30 — the same matrix).  Error types and messages follow the reference because callers and
tests match on them.
"""
from __future__ import annotations

import numpy as np

_SHAPE_4X4 = "4x4 homogeneous transformation matrix expected"
_SHAPE_3X3 = "3x3 matrix expected"
_SHAPE_T = "3-element translation vector expected"


def _compose(rotation: np.ndarray, translation: np.ndarray) -> np.ndarray:
    out = np.zeros((4, 4), dtype=float)
    out[3, 3] = 1.0
    out[0:3, 0:3] = rotation
    out[0:3, 3] = translation
    return out


class Transform3D:
    """Wrapper around one 4x4 homogeneous matrix (kept by reference, like the original)."""

    __slots__ = ("_Tmat",)

    def __init__(self, Tmat):
        if tuple(Tmat.shape) != (4, 4):
            raise ValueError(_SHAPE_4X4)
        self._Tmat = Tmat

    # -- constructors ------------------------------------------------------------------------
    @classmethod
    def from_rmat_t(cls, rmat=None, t=None) -> "Transform3D":
        rotation = np.eye(3, dtype=float) if rmat is None else rmat
        if tuple(rotation.shape) != (3, 3):
            raise ValueError(_SHAPE_3X3)
        translation = np.zeros(3, dtype=float) if t is None else np.asarray(t).ravel()
        if translation.size != 3:
            raise ValueError(_SHAPE_T)
        return cls(_compose(rotation, translation))

    @classmethod
    def identity(cls) -> "Transform3D":
        return cls(np.eye(4, dtype=float))

    # -- views -------------------------------------------------------------------------------
    Tmat = property(lambda self: self._Tmat)
    Rmat = property(lambda self: self._Tmat[0:3, 0:3])

    def _get_t(self):
        return self._Tmat[0:3, 3]

    def _set_t(self, value):
        self._Tmat[0:3, 3] = value

    t = property(_get_t, _set_t)

    # -- algebra -----------------------------------------------------------------------------
    def inv(self) -> "Transform3D":
        return type(self)(np.linalg.inv(self._Tmat))

    def __mul__(self, other) -> "Transform3D":
        if not isinstance(other, Transform3D):
            raise TypeError(f"Multiplication is only supported between {self.__class__} objects.")
        return type(self)(self._Tmat @ other._Tmat)

    __matmul__ = __mul__

    def __str__(self) -> str:
        return "Homogeneous transformation(\n%s)" % (self._Tmat,)
