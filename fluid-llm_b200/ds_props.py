"""The boundary type between the datasets and the model-side patch ops.

Same constructor arguments and attribute names as `/root/reference/src/dataloader/ds_props.py:4-25` (`DSProps`),
so `get_data_loader` (src/utils_model.py:41-45) and every consumer of `ds_props.*` keep working.  The derived
sizes are computed properties here instead of `__post_init__` assignments.
"""
from __future__ import annotations


class DSProps:
    __slots__ = ("Nx_patch", "Ny_patch", "patch_size", "seq_len", "channel", "downscale")

    def __init__(self, Nx_patch: int, Ny_patch: int, patch_size, seq_len: int, channel: int = 3, downscale: int = 1, **_derived):
        # **_derived: the reference's dataclass also accepts (and then overwrites) the derived fields
        self.Nx_patch, self.Ny_patch = int(Nx_patch), int(Ny_patch)
        self.patch_size = (int(patch_size[0]), int(patch_size[1]))
        self.seq_len, self.channel, self.downscale = seq_len, int(channel), int(downscale)

    @property
    def N_patch(self) -> int:                     # total number of patches per frame
        return self.Nx_patch * self.Ny_patch

    @property
    def input_tot_size(self):                     # image size in pixels that the patches tile
        return self.Nx_patch * self.patch_size[0], self.Ny_patch * self.patch_size[1]

    @property
    def out_tot_size(self):
        x, y = self.input_tot_size
        return x // self.downscale, y // self.downscale

    @property
    def out_patch_size(self):
        return self.patch_size[0] // self.downscale, self.patch_size[1] // self.downscale

    def __repr__(self):
        return (f"DSProps(Nx_patch={self.Nx_patch}, Ny_patch={self.Ny_patch}, patch_size={self.patch_size}, "
                f"seq_len={self.seq_len}, channel={self.channel}, downscale={self.downscale})")
