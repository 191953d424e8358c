import os, sys
sys.path.insert(0, '/root/repo')
import torch, bench
import fluid_llm_b200
from fluid_llm_b200.field_path import AIRFOIL, CYLINDER
from fluid_llm_b200.mesh_utils import MeshPlan
def pairs(ent, n):
    P = ent.shape[0]
    blocks = ent.view(P // 32, 2, 4, 4, 3).permute(0, 2, 4, 1, 3).reshape(-1, 8)
    return blocks
def wavefronts(blocks, slot, n):
    # per row: distinct entries -> bank group -> max count per group
    s = torch.where(blocks < n, blocks, torch.full_like(blocks, -1))
    srt, _ = torch.sort(s, dim=1)
    first = torch.ones_like(srt, dtype=torch.bool); first[:, 1:] = srt[:, 1:] != srt[:, :-1]
    valid = first & (srt >= 0)
    grp = torch.where(valid, slot[srt.clamp(min=0)] % 8, torch.full_like(srt, 8))
    cnt = torch.zeros((blocks.shape[0], 9), dtype=torch.int64, device=blocks.device).scatter_add_(1, grp, torch.ones_like(grp))
    wf = cnt[:, :8].max(dim=1).values.clamp(min=1)
    return wf.double().mean().item(), valid.sum(dim=1).double().mean().item()
for name in ("airfoil", "cylinder", "eagle"):
    w = bench.WORKLOADS[name]
    pers = AIRFOIL if w["personality"] == "airfoil" else CYLINDER
    meshes, _ = bench.make_inputs(dict(w, T=2, n_traj=1), 0)
    pos, cells, _ = meshes[0]
    plan = MeshPlan(pos, cells, 238)
    tab = plan.patch_table((16, 16), pers.crop_patches, pers.flip_y)
    ps = (plan.n_nodes + 3) // 4 * 4
    inside = tab.idx[:, 3] >= 0
    ent = torch.where(inside.unsqueeze(1), tab.idx[:, :3].long(), torch.full((1, 1), ps, dtype=torch.int64, device="cuda")).contiguous()
    blocks = pairs(ent, ps)
    morton = plan.node_slot_d.long()
    col_slot, _ = tab.coloured_slots(ps)
    ident = torch.arange(ps, device="cuda")
    print(name, "wavefronts per quarter-warp gather (distinct nodes %.2f):" % wavefronts(blocks, ident, ps)[1],
          "identity %.3f" % wavefronts(blocks, ident, ps)[0], "morton %.3f" % wavefronts(blocks, morton, ps)[0],
          "coloured %.3f" % wavefronts(blocks, col_slot.long(), ps)[0])
