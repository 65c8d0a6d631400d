"""Mirror of lib/transforms/transforms.py:10-66 (4x4 homogeneous transform wrapper).

Same constructor checks, properties and operators; ``from_rmat_t`` composes the matrix
directly instead of calling transforms3d.affines.compose (transforms.py:30, unit zooms).
"""
from __future__ import annotations

from typing import Optional

import numpy as np


class Transform3D:
    def __init__(self, Tmat):
        if (4, 4) != Tmat.shape:
            raise ValueError("4x4 homogeneous transformation matrix expected")
        self._Tmat = Tmat

    @classmethod
    def from_rmat_t(cls, rmat: Optional[np.ndarray] = None, t: Optional[np.ndarray] = None) -> "Transform3D":
        if rmat is None:
            rmat = np.eye(3, dtype=float)
        if rmat.shape != (3, 3):
            raise ValueError("3x3 matrix expected")
        if t is None:
            t = np.zeros((3,), dtype=float)
        t = np.asarray(t).reshape(-1)
        if t.size != 3:
            raise ValueError("3-element translation vector expected")
        T = np.eye(4, dtype=float)
        T[:3, :3] = rmat
        T[:3, 3] = t
        return cls(T)

    @classmethod
    def identity(cls) -> "Transform3D":
        return cls.from_rmat_t(np.eye(3, dtype=float), np.zeros((3,), dtype=float))

    @property
    def Tmat(self):
        return self._Tmat

    @property
    def t(self):
        return self._Tmat[:3, 3]

    @t.setter
    def t(self, value):
        self._Tmat[:3, 3] = value

    @property
    def Rmat(self):
        return self._Tmat[:3, :3]

    def inv(self):
        return self.__class__(np.linalg.inv(self.Tmat))

    def __mul__(self, other: "Transform3D") -> "Transform3D":
        if isinstance(other, Transform3D):
            return self.__class__(self.Tmat @ other.Tmat)
        raise TypeError(f"Multiplication is only supported between {self.__class__} objects.")

    def __matmul__(self, other: "Transform3D") -> "Transform3D":
        return self * other

    def __str__(self) -> str:
        return f"Homogeneous transformation(\n{self.Tmat})"
