"""Mirror of `/root/reference/src/dataloader/ds_props.py:4-25` (the boundary type between the
datasets and the model-side patch ops)."""
from dataclasses import dataclass


@dataclass
class DSProps:
    Nx_patch: int
    Ny_patch: int
    patch_size: tuple
    seq_len: int
    channel: int = 3
    downscale: int = 1
    input_tot_size: tuple = None
    out_tot_size: tuple = None
    tot_py: int = None
    N_patch: int = None
    out_patch_size: tuple = None

    def __post_init__(self):
        px, py = self.patch_size
        self.input_tot_size = (self.Nx_patch * px, self.Ny_patch * py)
        self.out_tot_size = (self.Nx_patch * px // self.downscale, self.Ny_patch * py // self.downscale)
        self.N_patch = self.Nx_patch * self.Ny_patch
        self.out_patch_size = (px // self.downscale, py // self.downscale)
