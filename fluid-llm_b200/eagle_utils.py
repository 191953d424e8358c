"""Mesh-baseline evaluator on the GPU data path: N-RMSE of node states compared on the regular grid.

Mirror of `/root/reference/eagle/eagle_utils.py:60-130` (`aux_calc_n_rmse`, `calc_n_rmse`, `get_nrmse`): the
reference grids true and predicted node states with 2 x 3 x seq_len separate `to_grid` calls on the CPU; here all
6 x seq_len fields go through one `fl_to_grid` launch and the error reduction stays on the device.
"""
from __future__ import annotations

import torch

from .mesh_utils import get_mesh_interpolation, to_grid


def aux_calc_n_rmse(preds: torch.Tensor, target: torch.Tensor, bc_mask: torch.Tensor):
    """eagle_utils.py:60-68."""
    error = (preds - target) * (~bc_mask)
    return torch.sqrt(error.pow(2).mean(dim=(-1, -2, -3)))


def calc_n_rmse(preds: torch.Tensor, target: torch.Tensor, bc_mask: torch.Tensor):
    """eagle_utils.py:71-86: (bs, seq_len, channel, px, py) -> velocity N-RMSE + pressure N-RMSE per step."""
    v = aux_calc_n_rmse(preds[:, :, :2], target[:, :, :2], bc_mask[:, :, :2])
    p = aux_calc_n_rmse(preds[:, :, 2:], target[:, :, 2:], bc_mask[:, :, 2:])
    return v + p


def get_nrmse(true_states, pred_states, mesh_pos, faces):
    """eagle_utils.py:89-130: states (bs, seq_len, N, 3) on the mesh of batch element 0 -> N-RMSE (1, seq_len)."""
    bs, seq_len, n_points, C = true_states.shape
    pos = mesh_pos[0, 0].detach().cpu().numpy()
    tri = faces[0, 0].detach().cpu().numpy()
    triang, tri_index, grid_x, grid_y = get_mesh_interpolation(pos, tri)
    dev = triang.device
    # (2, seq_len, 3, N) -> one launch over 6 * seq_len scalar fields
    fields = torch.stack([true_states[0], pred_states[0]]).to(dev, torch.float32).permute(0, 1, 3, 2).reshape(-1, n_points)
    data, mask = to_grid(fields.contiguous(), grid_x, grid_y, triang, tri_index)
    nx, ny = data.shape[-2:]
    data = data.view(2, seq_len, 3, nx, ny)
    mask = mask.view(2, seq_len, 3, nx, ny)[0, -1, 2]          # the reference keeps the last true-state pressure mask (:101,115)
    mask = mask.view(1, 1, 1, nx, ny).repeat(1, seq_len, 3, 1, 1)
    return calc_n_rmse(data[1].unsqueeze(0), data[0].unsqueeze(0), mask)
