"""Input format of the path: CSV -> Observation[] -- mirror of get_observations (src/hmm_utils.jl:19-31)."""
from __future__ import annotations

import csv
from typing import List, Optional, Sequence

import numpy as np

from .structs import Observation


def get_observations(source, time_col: int = 1, type_col: int = 0, val_seq: Optional[Sequence[int]] = None) -> List[Observation]:
    """get_observations(df; time_col=1, type_col=0, val_seq=2:size(df,2)) / get_observations(fpath).
    `source` is a file path or a 2-d array; column indices are 1-based like the reference."""
    if isinstance(source, str):
        with open(source, newline="") as f:
            rows = [r for r in csv.reader(f) if r]
        try:
            [float(v) for v in rows[0]]
        except ValueError:
            rows = rows[1:]  # header line
        data = np.asarray([[float(v) for v in r] for r in rows], dtype=np.float64)
    else:
        data = np.atleast_2d(np.asarray(source, dtype=np.float64))
    ncol = data.shape[1]
    if val_seq is None:
        val_seq = range(2, ncol + 1)
    obs = []
    for i in range(data.shape[0]):
        obs_type = 1 if type_col == 0 else int(data[i, type_col - 1])
        vals = np.asarray([int(data[i, j - 1]) for j in val_seq], dtype=np.int64)
        obs.append(Observation(float(data[i, time_col - 1]), obs_type, 1.0, vals))
    obs.sort(key=lambda o: o.time)
    return obs
