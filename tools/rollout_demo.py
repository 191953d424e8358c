"""BASELINE config 4: autoregressive rollout feeding OPT-125m-architecture patch tokens on one B200.

The data path of `src/models/model.py:168-233` (_generate / gen_seq) on this repo's kernels -- tcgen05 patch
embedding, fused rollout step (img_to_patch + mask-zero + add), patch_to_img at the end -- around a STOCK PyTorch
backbone (`transformers.OPTModel`, random init: no weights are available offline; architecture = opt-125m defaults)
and a stand-in linear token decoder (the reference's MLP+GNN decoder is out of scope).  Reports time per predicted
step split into backbone vs data path.     usage: python tools/rollout_demo.py [batch] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fluid_llm_b200 import synth
from fluid_llm_b200.ds_props import DSProps
from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
from fluid_llm_b200.mesh_utils import MeshPlan
from fluid_llm_b200.patch_embed import PatchEmbedder, RolloutEmbedCache
from fluid_llm_b200.simple_dataloader import position_ids
from fluid_llm_b200.utils_model import patch_to_img, rollout_step


def main():
    from transformers import OPTConfig, OPTModel
    bs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 251           # src/inference.py:87
    ctx = 10                                                          # configs/training1.yaml: autoreg_seq_len
    dev = torch.device("cuda")
    torch.manual_seed(0)
    # initial state + boundary mask from the GPU data path (Cylinder-shaped synthetic trajectories)
    states, masks = [], []
    for b in range(bs):
        tr = synth.make_trajectory("cylinder", 4, mesh_seed=b, field_seed=10 + b)
        plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
        s, m, tab = interp_patchify(DeviceTrajectory(tr["velocity"], tr["pressure"], plan), 0, 1, 1, (16, 16), CYLINDER)
        states.append(s)
        masks.append(m)
    props = DSProps(tab.n_bx, tab.n_by, (16, 16), ctx)
    L = props.N_patch
    state0 = torch.stack(states)                                       # (bs, 1, L, 3, 16, 16)
    bc = torch.stack(masks).bool().unsqueeze(3).expand(bs, 1, L, 3, 16, 16).contiguous()
    backbone = OPTModel(OPTConfig()).to(dev, torch.bfloat16).eval()    # opt-125m architecture: 768 hidden, 12 layers, 12 heads
    d = backbone.config.hidden_size
    embed = PatchEmbedder(torch.randn(512, 768) * 768 ** -0.5, torch.zeros(512), torch.randn(d, 512) * 512 ** -0.5, torch.zeros(d),
                          torch.randn(L, d) * 0.02, torch.randn(L, d) * 0.02, torch.randn(n_steps + ctx + 1, d) * 0.02)
    bos = torch.randn(1, 1, d, device=dev, dtype=torch.bfloat16)
    decoder = torch.nn.Linear(d, 768).to(dev, torch.bfloat16)          # stand-in for PatchDecoder (out of scope)
    pos_all = position_ids(n_steps + ctx + 1, props.Nx_patch, props.Ny_patch).to(dev)      # (T, L, 3)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_path = t_bb = 0.0
    n_timed = 0
    cached = "--reembed" not in sys.argv          # default: RolloutEmbedCache (one new state embedded per step)
    buf = [state0]
    g_embed = None
    cache = RolloutEmbedCache(embed, bs, L, ctx, graphs=True)
    tok = state0
    all_states = [state0]
    with torch.no_grad():
        for step in range(1, n_steps + 1):
            e0, e1, e2, e3 = ev(), ev(), ev(), ev()
            c = min(len(buf), ctx)
            ids = pos_all[:c].unsqueeze(0).expand(bs, c, L, 3)                     # time ids re-based to 0 (model.py:196-199)
            e0.record()
            if cached:
                steady = step > ctx + 1              # ring full, ids unchanged from the previous step, tokens already in the cache's buffer
                emb = cache.step(None if steady else tok, None if steady else ids)   # only the newest state goes through the GEMMs
            elif c == ctx:                                                         # steady state: fixed shape -> CUDA-graph form
                seq = torch.cat(buf[-ctx:], dim=1)                                 # (bs, c, L, 3, 16, 16)
                if g_embed is None:
                    g_embed = embed.graphed(bs * ctx * L, True, torch.float32)
                    e0.record()
                emb = g_embed(seq, ids).view(bs, c * L, d)
            else:
                emb = embed(torch.cat(buf[-ctx:], dim=1), ids, validate_ids=False).view(bs, c * L, d)      # tcgen05 patch embedding
            e1.record()
            x = torch.cat([bos.expand(bs, 1, d), emb.to(torch.bfloat16)], dim=1)
            h = backbone(inputs_embeds=x).last_hidden_state[:, -L:]                # stock PyTorch backbone
            pred = decoder(h).float().view(bs, 1, L, 3, 16, 16) * 0.05             # diff_scale_factor
            e2.record()
            pred_img = patch_to_img(pred, props)                                   # decoder output is image-shaped in the reference
            nxt, _, tok = rollout_step(buf[-1], pred_img, bc, props, tokens_out=cache.token_buffer)   # model.py:164,206,210 fused (+ bf16 tokens)
            e3.record()
            torch.cuda.synchronize()
            if step > 2 * ctx + 2:                  # steady state: the ring is full and every CUDA graph has been captured
                t_path += e0.elapsed_time(e1) + e2.elapsed_time(e3)
                t_bb += e1.elapsed_time(e2)
                n_timed += 1
            buf.append(nxt)
            buf = buf[-ctx:]
            all_states.append(nxt)
        imgs = patch_to_img(torch.cat(all_states, dim=1), props)                   # model.py:231
    tokens = min(ctx, n_steps) * L + 1
    print(f"rollout: batch {bs}, {n_steps} steps, context {ctx} states x {L} patches + BOS = {tokens} tokens, output {tuple(imgs.shape)}")
    print("embedding: " + ("RolloutEmbedCache (new state only + positional add over the ring)" if cached else "whole context re-embedded (graphed)"))
    n_timed = max(n_timed, 1)
    print(f"per predicted step (steady state, {n_timed} steps): data path {t_path / n_timed * 1e3:.1f} us (embed + unpatchify + fused step), "
          f"backbone+decoder (stock PyTorch bf16) {t_bb / n_timed * 1e3:.1f} us -> data path = "
          f"{100 * t_path / (t_path + t_bb):.1f} % of the step")
    assert torch.isfinite(imgs).all()


if __name__ == "__main__":
    main()
