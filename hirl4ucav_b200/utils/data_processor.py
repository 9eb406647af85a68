"""Expert-data CSV I/O (reference hirl/utils/data_processor.py:5-17): the on-disk format a user
brings.  Two CSV rows -- states, actions -- whose cells are ``str(np.ndarray)`` of one sample each.
Host-side file parsing only; nothing numerical happens here."""
import csv

import numpy as np
import pandas as pd


def _parse_row(cells):
    return np.array([np.fromstring(item[1:-1], sep=" ") for item in cells])


def read_data(data_dir):
    expert_data = pd.read_csv(data_dir, header=None)
    return _parse_row(expert_data.iloc[0].to_numpy()), _parse_row(expert_data.iloc[1].to_numpy())


def write_data(npstate, npaction, data_dir):
    with open(data_dir, "w", newline="") as f:
        csv.writer(f).writerows([npstate, npaction])


def up_sample(BCActions):
    return np.where(BCActions[:, 3] == 1)[0]
