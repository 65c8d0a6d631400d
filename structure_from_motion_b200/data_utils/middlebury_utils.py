"""Mirror of lib/data_utils/middlebury_utils.py:15-54: reads camera parameters of the Middlebury multi-view
datasets (https://vision.middlebury.edu/mview/data/).  Text parsing only — nothing here is on a compute path."""
import re
from pathlib import Path

import numpy as np

from ..transforms.transforms import Transform3D


def load_camera_k_r_t(par_filepath: Path, file_index: int):
    """(K [3,3], world->camera Transform3D) of image ``file_index`` from a ``*_par.txt`` file whose first line is the
    number of entries and whose other lines read ``name<index>.png k11..k33 r11..r33 t1 t2 t3``.
    ValueError when the index exceeds the entry count or is not listed, RuntimeError for an undecodable file name."""
    with Path(par_filepath).open("rt") as par_file:
        num_entries = int(par_file.readline())
        if file_index > num_entries:
            raise ValueError(f"There are {num_entries} entries in {par_filepath}, requested entry no. {file_index}.")
        for line in par_file:
            parts = line.split(" ")
            match = re.match(r"^.+?([\d]+)\.png$", parts[0])
            if match is None:
                raise RuntimeError(f"Could not decode filename {parts[0]}.")
            if int(match[1]) == file_index:
                values = [float(p) for p in parts[1:22]]
                k = np.array(values[0:9]).reshape(3, 3)
                r = np.array(values[9:18]).reshape(3, 3)
                t = np.array(values[18:21]).reshape(3, 1)
                return k, Transform3D.from_rmat_t(r, t)
        raise ValueError(f"Could not find matching entry for file index {file_index}")
