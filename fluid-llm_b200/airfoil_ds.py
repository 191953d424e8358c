"""`AirfoilDataset` on the GPU data path.

Host-side mirror of `/root/reference/src/dataloader/airfoil_ds.py:24-257`: node crop to
x in (-0.5, 2), y in (-0.75, 0.75) with face filter and renumbering (:164-183), y flip after
padding (:80), outer ring of patches dropped (:132-133, N_*_patch - 2 at :54), mask-aware
normalisation (:216-244), natural-sorted file list (:44).
"""
from __future__ import annotations

import re

import numpy as np

from .field_path import AIRFOIL
from .simple_dataloader import _GpuFieldDataset


def _natsorted(seq):
    return sorted(seq, key=lambda s: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)])


def crop_airfoil_mesh(pos, faces):
    """airfoil_ds.py:164-183 -> (node_mask, cropped pos, renumbered faces)."""
    pos = np.asarray(pos)
    faces = np.asarray(faces)
    mask = (pos[:, 0] > -.5) & (pos[:, 0] < 2) & (pos[:, 1] > -.75) & (pos[:, 1] < 0.75)
    wanted_nodes = np.nonzero(mask)[0]
    all_nodes = np.zeros(len(mask), dtype=np.int64)
    all_nodes[mask] = np.arange(len(wanted_nodes), dtype=np.int64)
    face_mask = mask[faces].all(axis=1)          # == np.isin(faces, wanted_nodes).all(axis=1)
    return mask, pos[mask], all_nodes[faces[face_mask]]


class AirfoilDataset(_GpuFieldDataset):
    """Load a sequence of timesteps of one Airfoil trajectory (airfoil_ds.py:24)."""
    personality = AIRFOIL
    ingest_airfoil_crop = True       # the ingest workers apply the node crop of _prepare_mesh

    def _list_files(self):
        return _natsorted(self._one_per_stem())

    def _prepare_mesh(self, save_data):
        mask, pos, faces = crop_airfoil_mesh(save_data['mesh_pos'], save_data['cells'])
        vel = np.ascontiguousarray(save_data['velocity'][:, mask])
        prs = np.ascontiguousarray(save_data['pressure'][:, mask])
        return pos, faces, vel, prs
