"""The data-path side of the autoregressive rollout, behind the reference model's own method names.

Mirror of `/root/reference/src/models/model.py:154-233` (`_gen_step`, `_generate`, `gen_seq`): the sliding context buffer,
the re-based time ids, `diffs = img_to_patch(pred); diffs[mask] = 0; next = last + diffs`, and the final unpatchify.  The
backbone stays whatever the caller passes as `forward_see_init(states, position_ids) -> (bs, seq_len, 3, tot_px, tot_py)`
(the reference's OPT / LoRA stack, out of scope here); everything between two backbone calls is one fused kernel
(`fl_rollout_step`) plus views.  Outputs are bit-identical to the reference's loop (tests/golden/ref_callers.npz was
produced by the reference's unmodified methods).
"""
from __future__ import annotations

from collections import deque

import torch

from .ds_props import DSProps
from .utils_model import patch_to_img, rollout_step


class RolloutGlue:
    """Holds what the reference's methods read from `self`: `ds_props`, `max_ctx_len` and the backbone call."""

    def __init__(self, forward_see_init, ds_props: DSProps, max_ctx_len: int):
        self.forward_see_init = forward_see_init
        self.ds_props = ds_props
        self.max_ctx_len = int(max_ctx_len)

    @torch.no_grad()
    def _generate(self, init_states, bc_mask, position_ids, N_steps):
        """model.py:168-216.  init_states (bs, init_len, N_patch, 3, px, py) -> (all_states (bs, init_len + N_steps, ...),
        all_diffs (bs, N_steps, ...))."""
        bs, init_len, N_patch, channel, px, py = init_states.shape
        all_states = torch.empty((bs, init_len + N_steps, N_patch, channel, px, py), dtype=init_states.dtype, device=init_states.device)
        all_diffs = torch.empty((bs, N_steps, N_patch, channel, px, py), dtype=init_states.dtype, device=init_states.device)
        all_states[:, :init_len] = init_states
        input_buff = deque(maxlen=self.max_ctx_len)
        for t in range(init_len):
            input_buff.append(all_states[:, t:t + 1])
        for pred_step in range(init_len, init_len + N_steps):
            seq_len = len(input_buff)
            start_pos = pred_step - seq_len
            seq_pos_ids = position_ids[:, start_pos:pred_step].clone()
            seq_pos_ids[:, :, :, 2] -= seq_pos_ids[:, :, :, 2].min()          # first state of the context is t = 0 (:196-199)
            mask = bc_mask[:, pred_step - 1: pred_step]
            s = torch.cat(list(input_buff), dim=1)
            pred = self.forward_see_init(s, seq_pos_ids)[:, -1:]              # _gen_step (:154-166) without the re-patchify ...
            nxt, diffs = rollout_step(input_buff[-1], pred, mask, self.ds_props)   # ... which is fused with the mask and the add
            all_diffs[:, pred_step - init_len: pred_step - init_len + 1] = diffs
            all_states[:, pred_step: pred_step + 1] = nxt
            input_buff.append(all_states[:, pred_step: pred_step + 1])
        return all_states, all_diffs

    @torch.no_grad()
    def gen_seq(self, batch_data, pred_steps, start_state=1):
        """model.py:218-233 -> (all_states, all_diffs) as images (bs, seq_len, 3, tot_px, tot_py)."""
        states, _, _, bc_mask, position_ids = batch_data
        bs, seq_len, N_patch, channel, px, py = states.shape
        assert pred_steps + start_state - 1 <= seq_len, \
            f'Prediction steps ({pred_steps}) + start state ({start_state}) must be less than total sequence length {seq_len}!'
        all_states, all_diffs = self._generate(states[:, :start_state], bc_mask, position_ids, pred_steps)
        return patch_to_img(all_states, self.ds_props), patch_to_img(all_diffs, self.ds_props)
