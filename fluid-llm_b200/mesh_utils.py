"""Mesh -> regular grid on the GPU, behind the reference's own call signatures.

Mirror of `/root/reference/src/dataloader/mesh_utils.py:64-106` (`grid_pos`, `to_grid`,
`get_mesh_interpolation`).  The matplotlib objects the reference passes around (`triang`, the
trifinder) become a `MeshPlan`: device-resident node positions, corrected triangles, triangle id
per grid cell and the static (vertex ids, barycentric weights) table the per-step kernel reads.
All arithmetic happens in libfluidgrid.so (csrc/fl_locate.cu, fl_interp.cu); this file only
moves arrays and keeps the reference's argument checks and error behaviour.
"""
from __future__ import annotations

import contextlib
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import FL_FLIP_Y, FL_NO_PAD, check, load, ptr, stream_ptr

from ._plan_host import (F32, F64, _grid_axes, _grid_axis, _grid_shape, default_numpy_semantics, morton_slots,  # noqa: F401
                         prepare_plan)


def grid_pos(x_min, x_max, y_min, y_max, grid_res, numpy_semantics=None):
    """mesh_utils.py:64-79 -> (grid_x, grid_y) float32 (nx, ny), index order [ix, iy]."""
    sem = numpy_semantics or default_numpy_semantics()
    ax, ay = _grid_axes(float(x_min), float(x_max), float(y_min), float(y_max), int(grid_res), sem)
    nx, ny = len(ax), len(ay)
    return (np.ascontiguousarray(np.broadcast_to(ax[:, None], (nx, ny))),
            np.ascontiguousarray(np.broadcast_to(ay[None, :], (nx, ny))))




def serpentine_patches(n_bx, n_by):
    """Patch ids l = bx * n_by + by along a serpentine through the patch grid (up column 0, down column 1, ...):
    consecutive runs of this order are spatially compact sets of patches."""
    l = np.arange(n_bx * n_by, dtype=np.int32).reshape(n_bx, n_by)
    l[1::2] = l[1::2, ::-1]
    return l.reshape(-1)


def tile_sizes(n_patches, tp):
    """Split n_patches into ceil(n_patches / tp) runs of nearly equal length (they differ by at most one patch)."""
    n_tiles = max(1, -(-n_patches // tp))
    base, extra = divmod(n_patches, n_tiles)
    return [base + 1] * extra + [base] * (n_tiles - extra)


def choose_tile_patches(n_patches, n_nodes, ppx):
    """Patches per tile for the tiled kernel (csrc/fl_tiled.cu): one patch per patch group of consumer warps, i.e.
    14 consumer warps / (ppx / 128) warps per patch with two producer warps, or 12 / (ppx / 128) with four producer
    warps for meshes with many nodes per output pixel (their staging needs more threads)."""
    wpp = max(1, ppx // 128)
    dense = n_nodes > 0.2 * n_patches * ppx
    return max(1, min((12 if dense else 14) // wpp, n_patches))


def colour_entries(ent, n_entries, tile_of, n_tiles, n_colours=8, rounds=24, seed=0):
    """Bank-group colouring of the tiles' node entries.  ent: int64 [P, 3] = entry (tile-local node) of every output pixel's
    three vertices, n_entries for pixels outside the mesh; pixels in output order, P a multiple of 32, patch rows of 16.
    A gather instruction's quarter-warp serves a 2 x 4 block of pixels (csrc/fl_tiled.cu: lane permutation) and reads the
    k-th vertex of each: entries read by the same quarter-warp instruction should get different colours (= slot mod 8 = bank
    group).  Jones-Plassmann style: in every round the uncoloured entries whose random priority beats all their uncoloured
    neighbours take, among the colours no neighbour holds, the one their TILE has used least (tile_of: tile of every entry), so
    that a tile's colours stay balanced and its slot list short.  -> int64 [n_entries] colours (entries that found no free
    colour keep the one their fewest neighbours hold)."""
    dev = ent.device
    P = ent.shape[0]
    blocks = ent.view(P // 32, 2, 4, 4, 3).permute(0, 2, 4, 1, 3).reshape(-1, 8)          # [block x k, 8 lanes]
    ii, jj = torch.triu_indices(8, 8, offset=1, device=dev)
    a, b = blocks[:, ii].reshape(-1), blocks[:, jj].reshape(-1)
    ok = (a != b) & (a < n_entries) & (b < n_entries)
    a, b = a[ok], b[ok]
    if a.numel():
        e = torch.unique(torch.minimum(a, b) * n_entries + torch.maximum(a, b))
        a, b = e // n_entries, e % n_entries
    src, dst = torch.cat([a, b]), torch.cat([b, a])                                           # both directions
    g = torch.Generator(device=dev).manual_seed(seed)
    prio = torch.rand(n_entries, generator=g, device=dev, dtype=torch.float64)
    colour = torch.full((n_entries,), -1, dtype=torch.int64, device=dev)
    used = torch.zeros((n_entries, n_colours), dtype=torch.int32, device=dev)               # neighbours holding each colour
    load = torch.zeros((n_tiles, n_colours), dtype=torch.int64, device=dev)
    for _ in range(rounds):
        un = colour < 0
        if not bool(un.any()):
            break
        # highest priority among uncoloured neighbours
        nb = torch.where(un[src], prio[src], torch.full((1,), -1.0, dtype=torch.float64, device=dev))
        best = torch.full((n_entries,), -1.0, dtype=torch.float64, device=dev).scatter_reduce(0, dst, nb, "amax", include_self=True)
        pick = un & (prio > best)
        idx = torch.nonzero(pick).squeeze(1)
        if idx.numel() == 0:
            break
        # free colours first (fewest neighbours holding it), ties to the least-loaded colour
        score = used[idx].long() * (1 << 40) + load[tile_of[idx]]
        c = torch.argmin(score, dim=1)
        colour[idx] = c
        load.index_put_((tile_of[idx], c), torch.ones(idx.numel(), dtype=torch.int64, device=dev), accumulate=True)
        m = pick[src]
        used.index_put_((dst[m], colour[src[m]]), torch.ones(int(m.sum().item()), dtype=torch.int32, device=dev), accumulate=True)
    rest = torch.nonzero(colour < 0).squeeze(1)
    if rest.numel():
        colour[rest] = torch.argmin(used[rest].long() * (1 << 40) + load[tile_of[rest]], dim=1)
    return colour


class TilePlan:
    """The patch table split into tiles for csrc/fl_tiled.cu (FlTraj::d_idx_tile, d_tile_*)."""

    def __init__(self, idx_tile, tile_nodes, tile_desc, tile_patches, tile_quads, tile_qslots, n_tiles, max_tile_nodes, tp):
        self.idx_tile, self.tile_nodes, self.tile_desc, self.tile_patches = idx_tile, tile_nodes, tile_desc, tile_patches
        self.tile_quads, self.tile_qslots = tile_quads, tile_qslots
        self.n_tiles, self.max_tile_nodes, self.tp = n_tiles, max_tile_nodes, tp


class PatchTable:
    """The static table re-ordered into output-pixel order for one dataset personality.
    `idx_slot` is the same table with node ids replaced by the plan's shared-memory slots."""

    def __init__(self, idx, w, n_bx, n_by, px, py, idx_slot=None, n_nodes=0, node_rank=None):
        self.idx, self.w, self.n_bx, self.n_by, self.px, self.py = idx, w, n_bx, n_by, px, py
        self.idx_slot = idx_slot
        self.node_rank = node_rank          # rank of every node in a spatial (Morton) sort: orders the slots inside a tile
        self.n_patches = n_bx * n_by
        self.n_nodes = n_nodes
        self._tile_plans = {}

    def idx_slot16(self, n_slots):
        """The slot form of the table in 8-byte records (FlTraj::idx_slot_format = 1, fl_pack_idx16): the staged kernel reads
        each lane's table records once per work item, and a third fewer bytes there are 2 - 5 % of its launch.
        n_slots = the padded node count (slots are below it); <= 65536."""
        key = ("idx16", int(n_slots))
        got = self._tile_plans.get(key)
        if got is None:
            if self.idx_slot is None:
                raise ValueError("idx_slot16: the table has no slot form")
            got = torch.empty((self.idx_slot.shape[0], 2), dtype=torch.int32, device=self.idx_slot.device)
            with torch.cuda.device(got.device):
                check(load().fl_pack_idx16(ptr(self.idx_slot), self.idx_slot.shape[0], int(n_slots), ptr(got), stream_ptr()), "fl_pack_idx16")
            self._tile_plans[key] = got
        return got

    def coloured_slots(self, n_slots):
        """Shared-memory slots of the whole mesh's nodes for the staged kernel, by bank-group colouring instead of the plan's
        Morton order: -> (node_slot int32 [n_slots], idx_slot int32 [P, 4]).  Nodes that one quarter-warp gather reads together
        get different slot mod 8 where eight colours allow it; colour c owns the slots c, c + 8, ... below n_slots, so the
        records still fit the frame pitch.  n_slots = the padded node count (FlTraj::prs_stride)."""
        key = ("coloured", int(n_slots))
        got = self._tile_plans.get(key)
        if got is not None:
            return got
        dev = self.idx.device
        P = self.idx.shape[0]
        if P % 32 or self.py != 16 or n_slots < self.n_nodes:
            raise ValueError("coloured_slots: needs 16-pixel patch rows and n_slots >= n_nodes")
        inside = self.idx[:, 3] >= 0
        ent = torch.where(inside.unsqueeze(1), self.idx[:, :3].long(), torch.full((1, 1), n_slots, dtype=torch.int64, device=dev))
        zero = torch.zeros(n_slots, dtype=torch.int64, device=dev)
        col = colour_entries(ent.contiguous(), n_slots, zero, 1)
        rank = self.node_rank.long() if self.node_rank is not None else torch.arange(self.n_nodes, device=dev)
        order_key = torch.cat([rank, torch.arange(self.n_nodes, n_slots, device=dev)])         # Morton order inside a colour, pads last
        cap = (n_slots - torch.arange(8, device=dev) + 7) // 8                                     # slots congruent to c below n_slots

        def rank_in(c):
            o = torch.argsort(order_key, stable=True)
            o = o[torch.argsort(c[o], stable=True)]
            _, cnt = torch.unique_consecutive(c[o], return_counts=True)
            start = torch.cumsum(cnt, 0) - cnt
            r = torch.empty(n_slots, dtype=torch.int64, device=dev)
            r[o] = torch.arange(n_slots, device=dev) - torch.repeat_interleave(start, cnt)
            return r
        r_in = rank_in(col)
        excess = r_in >= cap[col]
        if bool(excess.any()):
            per = torch.bincount(col, minlength=8)
            free = (cap - per).clamp(min=0)
            ex_idx = torch.nonzero(excess).squeeze(1)
            j = torch.arange(ex_idx.numel(), device=dev)
            tgt = torch.searchsorted(torch.cumsum(free, 0), j, right=True).clamp(max=7)
            col = col.clone()
            col[ex_idx] = tgt
            r_in = rank_in(col)
        node_slot = (8 * r_in + col).to(torch.int32)
        idx_slot = torch.cat([torch.where(inside.unsqueeze(1), node_slot[self.idx[:, :3].clamp(min=0, max=n_slots - 1).long()], self.idx[:, :3]),
                              self.idx[:, 3:4]], dim=1).contiguous()
        self._tile_plans[key] = (node_slot.contiguous(), idx_slot)
        return self._tile_plans[key]

    def default_tile_patches(self):
        return choose_tile_patches(self.n_patches, self.n_nodes, self.px * self.py)

    def tile_plan(self, tp=None) -> TilePlan:
        """Tiles = runs of `tp` patches along the serpentine; per tile the sorted list of the nodes its pixels touch,
        and the table with node ids replaced by 16 * (position in that list).  Index bookkeeping only (one sort of
        3 * pixels keys on the device); built once per table and tile size."""
        tp = int(tp) if tp else self.default_tile_patches()
        plan = self._tile_plans.get(tp)
        if plan is not None:
            return plan
        dev = self.idx.device
        L, ppx, N = self.n_patches, self.px * self.py, max(int(self.n_nodes), 1)
        P_total = L * ppx
        colour_slots = os.environ.get("FLUIDGRID_COLOUR_SLOTS", "1") != "0"
        order = serpentine_patches(self.n_bx, self.n_by)
        sizes = tile_sizes(L, tp)
        n_tiles = len(sizes)
        patch_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        tile_of_patch = np.empty(L, dtype=np.int64)
        tile_of_patch[order] = np.repeat(np.arange(n_tiles), sizes)
        # inside a tile the patches are listed with ascending ids: runs of consecutive ids are contiguous blocks of a frame's
        # output, which csrc/fl_ring.cu stores with one bulk copy each
        order = np.concatenate([np.sort(order[patch_off[t]:patch_off[t + 1]]) for t in range(n_tiles)]).astype(order.dtype)
        with (torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()):
            tile_px = torch.from_numpy(tile_of_patch).to(dev).repeat_interleave(ppx)            # tile of every output pixel
            inside = self.idx[:, 3] >= 0
            sentinel = n_tiles * N
            key = torch.where(inside.unsqueeze(1), tile_px.unsqueeze(1) * N + self.idx[:, :3].long(),
                              torch.full((1, 1), sentinel, dtype=torch.int64, device=dev))
            uniq, inv = torch.unique(key.reshape(-1), return_inverse=True)                    # sorted: tile-major, node ids ascending
            n_valid = int((uniq < sentinel).sum().item())
            u_tile = uniq[:n_valid] // N
            node_u = uniq[:n_valid] % N
            counts = torch.bincount(u_tile, minlength=n_tiles)
            node_off = torch.cumsum(counts, 0) - counts
            # Slots inside a tile.  (1) A spatial (Morton) order of the nodes, not their ids: on a mesh numbered row by row the
            # nodes of neighbouring columns would sit a whole column apart, i.e. in the same bank group again and again.
            real = None
            if self.node_rank is not None and n_valid:
                sorder = torch.argsort(u_tile * N + self.node_rank[node_u].long(), stable=True)
                pos_of = torch.empty_like(sorder)
                pos_of[sorder] = torch.arange(n_valid, device=dev)
                pos_of = torch.cat([pos_of, torch.full((uniq.numel() - n_valid,), n_valid, dtype=pos_of.dtype, device=dev)])
                inv = pos_of[inv]
                u_tile, node_u = u_tile[sorder], node_u[sorder]
            # (2) On top of that, bank groups by colouring: entries that one quarter-warp gather reads together get different
            # slot mod 8; the r-th entry of a (tile, colour) in Morton order sits at slot 8 r + colour.  A tile's list is as long
            # as its fullest colour needs; unused slots repeat the tile's first node (staged, never gathered).
            if colour_slots and n_valid and P_total % 32 == 0 and self.py == 16:
                col = colour_entries(inv.reshape(-1, 3), n_valid, u_tile, n_tiles)

                def rank_in(key):            # position of every entry inside its run of equal keys, Morton order kept
                    o = torch.argsort(key, stable=True)
                    _, cnt = torch.unique_consecutive(key[o], return_counts=True)
                    start = torch.cumsum(cnt, 0) - cnt
                    r = torch.empty(n_valid, dtype=torch.int64, device=dev)
                    r[o] = torch.arange(n_valid, device=dev) - torch.repeat_interleave(start, cnt)
                    return r
                # every colour of a tile holds at most cap = ceil(n / 8) entries, so that the list is no longer than n + 7: the
                # entries beyond cap move, in order, into the free places of the tile's other colours (a few conflicts come back)
                cap = (counts + 7) // 8
                r_in = rank_in(u_tile * 8 + col)
                excess = r_in >= cap[u_tile]
                if bool(excess.any()):
                    per = torch.zeros((n_tiles, 8), dtype=torch.int64, device=dev)
                    per.index_put_((u_tile, col), torch.ones(n_valid, dtype=torch.int64, device=dev), accumulate=True)
                    free = (cap.unsqueeze(1) - per).clamp(min=0)                          # [tile, colour] places left
                    # the j-th excess entry of a tile takes the j-th free place of the tile (colours in order)
                    ex_idx = torch.nonzero(excess).squeeze(1)
                    j = rank_in(torch.where(excess, u_tile, torch.full_like(u_tile, n_tiles)))[ex_idx]
                    cum = torch.cumsum(free, dim=1)                                       # [tile, 8]
                    tgt = torch.searchsorted(cum[u_tile[ex_idx]].contiguous(), j.unsqueeze(1), right=True).squeeze(1).clamp(max=7)
                    col = col.clone()
                    col[ex_idx] = tgt
                    r_in = rank_in(u_tile * 8 + col)
                slot = 8 * r_in + col
                need = torch.zeros(n_tiles, dtype=torch.int64, device=dev).scatter_reduce(0, u_tile, slot + 1, "amax", include_self=True)
                if int(need.max().item()) <= int(1.25 * counts.max().item()) + 8:      # (badly unbalanced colours: keep plain Morton)
                    new_off = torch.cumsum(need, 0) - need
                    total = int(need.sum().item())
                    first_node = torch.zeros(n_tiles, dtype=torch.int64, device=dev)
                    first_node[u_tile.flip(0)] = node_u.flip(0)
                    nodes_full = torch.repeat_interleave(first_node, need)
                    gpos = new_off[u_tile] + slot
                    nodes_full[gpos] = node_u
                    place = torch.cat([gpos, torch.full((uniq.numel() - n_valid,), total, dtype=torch.int64, device=dev)])
                    inv = place[inv]
                    real = torch.zeros(total, dtype=torch.bool, device=dev)
                    real[gpos] = True
                    u_tile = torch.repeat_interleave(torch.arange(n_tiles, device=dev), need)
                    node_u = nodes_full
                    counts, node_off, n_valid = need, new_off, total
            local = (inv.reshape(-1, 3) - node_off[tile_px].unsqueeze(1)) * 16                  # byte offset of the node's record
            idx_tile = torch.cat([torch.where(inside.unsqueeze(1), local, torch.zeros_like(local)).to(torch.int32),
                                  self.idx[:, 3:4]], dim=1).contiguous()
            tile_nodes = node_u.to(torch.int32).contiguous()
            # per tile the quads (4 consecutive node ids) that hold its nodes, and per quad the slots of its nodes
            NQ = (N + 3) // 4 + 1
            if real is not None:                     # the quad lists describe the real nodes only (holes are never staged by quads)
                keep = torch.nonzero(real).squeeze(1)
                q_ut, q_nu, q_local = u_tile[keep], node_u[keep], (keep - node_off[u_tile[keep]])
            else:
                q_ut, q_nu, q_local = u_tile, node_u, torch.arange(n_valid, device=dev) - node_off[u_tile]
            qkey = q_ut * NQ + q_nu // 4
            uq, qinv = torch.unique(qkey, return_inverse=True)
            q_tile = uq // NQ
            tile_quads = (uq % NQ).to(torch.int32).contiguous()
            q_cnt = torch.bincount(q_tile, minlength=n_tiles)
            q_off = torch.cumsum(q_cnt, 0) - q_cnt
            q_max = torch.zeros(n_tiles, dtype=torch.int64, device=dev).scatter_reduce(0, q_tile, uq % NQ + 1, "amax", include_self=True)
            qslots = torch.full((max(int(uq.numel()), 1) * 4,), -1, dtype=torch.int32, device=dev)
            if n_valid:
                qslots[qinv * 4 + q_nu % 4] = q_local.to(torch.int32)
            if tile_quads.numel() == 0:
                tile_quads = torch.zeros(4, dtype=torch.int32, device=dev)
            if tile_nodes.numel() == 0:
                tile_nodes = torch.zeros(4, dtype=torch.int32, device=dev)
            desc = torch.stack([node_off, counts, torch.from_numpy(patch_off[:-1]).to(dev),
                                torch.from_numpy(np.asarray(sizes, dtype=np.int64)).to(dev),
                                q_off, q_cnt, q_max, torch.zeros_like(q_off)], dim=1).to(torch.int32).contiguous()
            tile_patches = torch.from_numpy(order).to(dev)
            max_nodes = int(counts.max().item()) if n_tiles else 0
        plan = TilePlan(idx_tile, tile_nodes, desc, tile_patches, tile_quads, qslots.view(-1, 4), n_tiles, max_nodes, tp)
        self._tile_plans[tp] = plan
        return plan


class MeshPlan:
    """Device-side stand-in for `matplotlib.tri.Triangulation` + its trifinder (mesh_utils.py:103-104).

    Opaque to callers, exactly like the reference's `triang`: it is only handed back to `to_grid`.
    `x`, `y`, `triangles` are kept (host) because the reference's interpolator checks `z` against
    `triangulation.x.shape` (src/_triinterpolate.py:37-39)."""

    def __init__(self, pos, faces, grid_res=238, numpy_semantics=None, device=None, allow_degenerate=False, sync=True,
                 prepared=None):
        """`allow_degenerate`: matplotlib's trapezoid-map trifinder is undefined on triangles of zero area (three colinear
        nodes: their edges overlap, its map builder raises or loops), and its plane fit takes a pseudo-inverse branch there
        (`calculate_plane_coefficients`); the data sets contain none.  By default such input raises ValueError like an
        invalid triangulation does upstream; with allow_degenerate=True the rule locator treats the triangle like any
        other and a grid point located in it gets the value of the triangle's first vertex (weights 0, 0).
        `sync=False` (the data sets' ingest path): nothing in the constructor waits for the GPU -- one pinned upload, the
        locate kernels with their status word left on the device; call `ready()` before trusting the tables (it re-locates
        with a larger workspace if the bin store overflowed, and says so).  `prepared`: the result of
        `_plan_host.prepare_plan` for the same arguments, computed elsewhere (pos / faces / grid_res are then not looked at)."""
        _lib.require_cuda()
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        # host side of the plan (validation as matplotlib's Triangulation does it, grid axes, Morton slots): NumPy only, so the
        # ingest workers can do it ahead of this process (`prepared`)
        h = prepared if prepared is not None else prepare_plan(pos, faces, grid_res, numpy_semantics, allow_degenerate)
        pos32, tri = h["pos32"], h["tri"]
        self.n_degenerate = int(h["n_degenerate"])
        self._pos32, self.triangles = pos32, tri
        self.n_nodes, self.n_cells = len(pos32), len(tri)
        self.numpy_semantics = h["numpy_semantics"]
        self.ax, self.ay = h["ax"], h["ay"]
        self.nx, self.ny = len(self.ax), len(self.ay)
        dev = self.device
        with torch.cuda.device(dev):
            self.n_padded = (self.n_nodes + 3) // 4 * 4
            slots = h["slots"]
            # one pinned pack, one upload: positions | triangles | grid axes | node slots (16-byte aligned pieces)
            parts = [pos32.reshape(-1).view(np.uint8), tri.reshape(-1).view(np.uint8), np.ascontiguousarray(self.ax).view(np.uint8),
                     np.ascontiguousarray(self.ay).view(np.uint8), slots.view(np.uint8)]
            offs, total = [], 0
            for a in parts:
                offs.append(total)
                total += (a.nbytes + 15) // 16 * 16
            pin = torch.empty(total, dtype=torch.uint8, pin_memory=True)
            pin_np = pin.numpy()
            for a, o in zip(parts, offs):
                pin_np[o:o + a.nbytes] = a
            pack = torch.empty(total, dtype=torch.uint8, device=dev)
            pack.copy_(pin, non_blocking=True)
            self._pin, self._pack = pin, pack           # the pinned source stays alive at least as long as the plan

            def piece(k, dtype, shape):
                return pack[offs[k]:offs[k] + parts[k].nbytes].view(dtype).view(shape)
            self.pos_d = piece(0, torch.float32, (self.n_nodes, 2))
            self.cells_d = piece(1, torch.int32, (self.n_cells, 3))
            self.ax_d = piece(2, torch.float32, (self.nx,))
            self.ay_d = piece(3, torch.float32, (self.ny,))
            self.node_slot_d = piece(4, torch.int32, (self.n_padded,))
            n = self.nx * self.ny
            self.tri_index_d = torch.empty((self.nx, self.ny), dtype=torch.int32, device=dev)
            self.cell_idx_d = torch.empty((n, 4), dtype=torch.int32, device=dev)
            self.cell_w_d = torch.empty((n, 2), dtype=torch.float64, device=dev)
            self._ws_bytes = int(load().fl_locate_workspace_bytes(self.n_nodes, self.n_cells))
            self._tri_index_host = None
            self._status = None
            if sync:
                self.locate()
            else:
                self._locate_async()
        self._tables = {}
        self._tri_index_host = None

    def _locate_async(self):
        """fl_locate_async on the current stream + the status word on its way to pinned memory; `ready()` reads it."""
        lib = load()
        with torch.cuda.device(self.device):
            self._ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
            st_d = torch.empty(2, dtype=torch.int32, device=self.device)
            check(lib.fl_locate_async(ptr(self.pos_d), ptr(self.cells_d), self.n_nodes, self.n_cells, ptr(self.ax_d), ptr(self.ay_d),
                                      self.nx, self.ny, ptr(self.tri_index_d), ptr(self.cell_idx_d), ptr(self.cell_w_d), ptr(self._ws),
                                      self._ws_bytes, ptr(st_d), stream_ptr()), "fl_locate_async")
            st_h = torch.empty(2, dtype=torch.int32, pin_memory=True)
            st_h.copy_(st_d, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._status = (st_h, ev, st_d)

    def ready(self):
        """True if the asynchronously built tables were fine (or the plan was built synchronously); False if the bin store had
        overflowed: the tables have then been rebuilt synchronously with a larger workspace and whatever was computed from
        the old ones must be redone."""
        if self._status is None:
            return True
        st_h, ev, _ = self._status
        ev.synchronize()
        self._status = None
        self._ws = None
        bad, over = int(st_h[0]), int(st_h[1])
        if bad == 0 and over == 0:
            return True
        self._ws_bytes += 4 * over + 256
        self._tables = {}
        self.locate()          # raises for bad node ids, as the synchronous constructor does
        return False

    def locate(self):
        """(Re)run the point-location kernels (fl_locate) on the resident mesh: fills tri_index and the static table.
        Called once by the constructor; bench.py calls it again to time the one-off step."""
        lib = load()
        with torch.cuda.device(self.device):
            for _ in range(2):
                ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
                rc = lib.fl_locate(ptr(self.pos_d), ptr(self.cells_d), self.n_nodes, self.n_cells, ptr(self.ax_d),
                                   ptr(self.ay_d), self.nx, self.ny, ptr(self.tri_index_d), ptr(self.cell_idx_d),
                                   ptr(self.cell_w_d), ptr(ws), self._ws_bytes, stream_ptr())
                if rc != -3:
                    break
                self._ws_bytes *= 8          # very uneven meshes: retry once with a larger bin-item store
            check(rc, "fl_locate")
        self._tri_index_host = None

    # -- reference-shaped views -------------------------------------------------------------
    @property
    def x(self):
        """float64 node x (matplotlib's `Triangulation.x`, which the interpolator checks `z` against)."""
        return self._pos32[:, 0].astype(F64)

    @property
    def y(self):
        return self._pos32[:, 1].astype(F64)

    @property
    def grid_x(self):
        return np.ascontiguousarray(np.broadcast_to(self.ax[:, None], (self.nx, self.ny)))

    @property
    def grid_y(self):
        return np.ascontiguousarray(np.broadcast_to(self.ay[None, :], (self.nx, self.ny)))

    @property
    def tri_index(self):
        """int32 (nx, ny) on the host, what `triang.get_trifinder()(grid_x, grid_y)` returns."""
        self.ready()
        if self._tri_index_host is None:
            self._tri_index_host = self.tri_index_d.cpu().numpy()
        return self._tri_index_host

    def patch_table(self, patch_size, crop_patches=0, flip_y=False, stride=None, pad=True) -> PatchTable:
        """The static table in output-pixel order.  `stride` (default: the patch size) and `pad` are the data sets' arguments of
        the same names (simple_dataloader.py:26-27,118-119,131): they only change which grid pixel an output pixel reads."""
        stride = patch_size if stride is None else stride
        key = (int(patch_size[0]), int(patch_size[1]), int(crop_patches), bool(flip_y), int(stride[0]), int(stride[1]), bool(pad))
        if key[4] < 1 or key[5] < 1:
            raise ValueError(f"stride must be positive, got {tuple(stride)}")
        tab = self._tables.get(key)
        if tab is None:
            px, py, sx, sy = key[0], key[1], key[4], key[5]
            lib = load()
            nbx, nby = ctypes.c_int(0), ctypes.c_int(0)
            flags = (FL_FLIP_Y if flip_y else 0) | (0 if pad else FL_NO_PAD)
            check(lib.fl_plan_patch_table(None, None, self.nx, self.ny, px, py, sx, sy, key[2], flags, None, None,
                                          ctypes.byref(nbx), ctypes.byref(nby), None, None, None), "fl_plan_patch_table")
            if nbx.value < 1 or nby.value < 1:
                raise ValueError(f"no patches left: grid {self.nx}x{self.ny}, patch {px}x{py}, stride {sx}x{sy}, crop {key[2]}, pad {pad}")
            total = nbx.value * nby.value * px * py
            with torch.cuda.device(self.device):
                idx = torch.empty((total, 4), dtype=torch.int32, device=self.device)
                idx_slot = torch.empty((total, 4), dtype=torch.int32, device=self.device)
                w = torch.empty((total, 2), dtype=torch.float64, device=self.device)
                # the table in output-pixel order, and the same with node ids as shared-memory slots, in one launch
                check(lib.fl_plan_patch_table(ptr(self.cell_idx_d), ptr(self.cell_w_d), self.nx, self.ny, px, py, sx, sy, key[2],
                                              flags, ptr(idx), ptr(w), ctypes.byref(nbx), ctypes.byref(nby),
                                              ptr(self.node_slot_d), ptr(idx_slot), stream_ptr()), "fl_plan_patch_table")
            tab = PatchTable(idx, w, nbx.value, nby.value, px, py, idx_slot, self.n_nodes, self.node_slot_d[:self.n_nodes])
            self._tables[key] = tab
        return tab


def get_mesh_interpolation(pos, faces, grid_res=238, numpy_semantics=None, device=None):
    """mesh_utils.py:94-106: -> (triang, tri_index int32 (nx,ny), grid_x, grid_y float32 (nx,ny)).

    `triang` is a MeshPlan; `tri_index`/`grid_*` are host arrays as in the reference."""
    if torch.is_tensor(pos):
        pos = pos.detach().cpu().numpy()
    if torch.is_tensor(faces):
        faces = faces.detach().cpu().numpy()
    plan = MeshPlan(pos, faces, grid_res, numpy_semantics, device)
    return plan, plan.tri_index, plan.grid_x, plan.grid_y


def to_grid(val, grid_x, grid_y, triang: MeshPlan, tri_index):
    """mesh_utils.py:82-91: one scalar node field -> (data float32 (nx,ny), mask bool (nx,ny)).

    NumPy in -> NumPy out (the reference's contract); a CUDA tensor in -> CUDA tensors out.
    `val` may also be (n_fields, N): all fields are gridded in one launch."""
    if not isinstance(triang, MeshPlan):
        raise TypeError("triang must be the MeshPlan returned by get_mesh_interpolation")
    as_torch = torch.is_tensor(val)
    v = val if as_torch else torch.from_numpy(np.ascontiguousarray(np.asarray(val), dtype=F32))
    single = v.dim() == 1
    if v.shape[-1:] != (triang.n_nodes,) or v.dim() > 2:
        raise ValueError("z array must have same length as triangulation x and y arrays")
    if tuple(np.shape(grid_x)) != tuple(np.shape(grid_y)):
        raise ValueError(f"x and y shall have same shapes. Given: {np.shape(grid_x)} and {np.shape(grid_y)}")
    if tuple(np.shape(tri_index)) != tuple(np.shape(grid_x)):
        raise ValueError("tri_index array is provided and shall have same shape as x and y. "
                         f"Given: {np.shape(tri_index)} and {np.shape(grid_x)}")
    if tuple(np.shape(grid_x)) != (triang.nx, triang.ny):
        raise ValueError(f"grid of shape {np.shape(grid_x)} does not belong to this mesh plan "
                         f"({triang.nx}, {triang.ny})")
    # the reference interpolates with the tri_index it is handed (_triinterpolate.py:265-267); the plan's static table was
    # built from the plan's own, so anything else is refused instead of being silently replaced
    if tri_index is not triang._tri_index_host:
        ti = tri_index.detach().cpu().numpy() if torch.is_tensor(tri_index) else np.asarray(tri_index)
        if not np.array_equal(ti, triang.tri_index):
            raise ValueError("tri_index differs from the one get_mesh_interpolation returned for this mesh plan")
    dev = triang.device
    with torch.cuda.device(dev):
        v = v.to(device=dev, dtype=torch.float32).reshape(-1, triang.n_nodes).contiguous()
        nf = v.shape[0]
        data = torch.empty((nf, triang.nx, triang.ny), dtype=torch.float32, device=dev)
        mask = torch.empty((nf, triang.nx, triang.ny), dtype=torch.uint8, device=dev)
        check(load().fl_to_grid(ptr(triang.cell_idx_d), ptr(triang.cell_w_d), triang.nx, triang.ny, ptr(v), nf,
                                triang.n_nodes, ptr(data), ptr(mask), stream_ptr()), "fl_to_grid")
    mask = mask.bool()
    if single:
        data, mask = data[0], mask[0]
    if as_torch:
        return data, mask
    return data.cpu().numpy(), mask.cpu().numpy()
