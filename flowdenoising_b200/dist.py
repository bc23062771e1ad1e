"""Multi-GPU slab sharding of the three passes (one process per GPU, torch.distributed / NCCL over NVLink).

Within a pass every output slice depends only on its 2r neighbour slices (src/flowdenoising.py:306-327), so each
pass shards its slice index into contiguous slabs with an r-slice PERIODIC halo (`% shape`, :312). Between passes
the slice orientation changes, so the volume is re-slabbed with an all-to-all:

    input       rank g holds the Z-slab            vol[zs_g:ze_g, :, :]
    Z pass      halo exchange (peer-to-peer send/recv of r slices each way), slab pass        -> A  [Zl][Y][X]
    re-slab     all-to-all; block g->h = A[:, (ys_h-r .. ye_h+r) mod Y, :]  (the sender owns every y of its z range,
                so the receiver's periodic halo travels with the block: no extra halo exchange) -> Ae [Z][Yl+2r][X]
    Y pass      slices are Ae[:, y, :]: slice stride X, row stride (Yl+2r)*X                    -> B  [Z][Yl][X]
    re-slab     all-to-all; block g->h = B[:, :, (xs_h-r .. xe_h+r) mod X], transposed while unpacking
                                                                                                -> Ce [Z][Xl+2r][Y]
    X pass      slices are Ce[:, x, :] (Z x Y images, y contiguous)                             -> D  [Z][Xl][Y]
    re-slab     all-to-all back to Z-slabs, transposed while unpacking                          -> out[Zl][Y][X]

Packing / transposing are the library's own kernels (fdn_copy3d, fdn_transpose_strided); the exchanges are NCCL
all-to-all / grouped send-recv (torch.distributed is plumbing only). Results are bit-identical for any number of
ranks because every slice is computed from identical inputs by the same kernels.

The compute object (`ops`) is a DeviceEngine; tests inject a CPU stand-in to exercise this host logic with the gloo
backend (tests/test_dist_gloo.py).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from ._lib import View


def split_range(n: int, parts: int, idx: int):
    """Balanced contiguous split of range(n) into `parts`; returns [start, end) of part idx."""
    base, rem = divmod(n, parts)
    start = idx * base + min(idx, rem)
    return start, start + base + (1 if idx < rem else 0)


class DistributedDenoiser:
    def __init__(self, ops, shape: Sequence[int], flow, exact: bool = True, group=None, chunk: Optional[int] = None):
        import torch.distributed as dist
        self.dist = dist
        self.ops = ops
        self.torch = ops.torch
        self.shape = tuple(int(s) for s in shape)
        self.flow = flow
        self.exact = exact
        self.chunk = chunk
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        Z, Y, X = self.shape
        if min(Z, Y, X) < self.world:
            raise ValueError("every axis must have at least one slice per rank")
        self.z_range = split_range(Z, self.world, self.rank)
        self.y_range = split_range(Y, self.world, self.rank)
        self.x_range = split_range(X, self.world, self.rank)
        self.backend = dist.get_backend(group)

    # ------------------------------------------------------------------ communication primitives
    def _all_to_all(self, send_flat, send_sizes, recv_flat, recv_sizes):
        """One flat float32 buffer per direction, split per peer (own block included)."""
        dist = self.dist
        if self.backend == "nccl":
            dist.all_to_all_single(recv_flat, send_flat, list(recv_sizes), list(send_sizes), group=self.group)
            return
        # gloo (CPU tests) has no all-to-all: the same exchange as grouped send/recv
        so = np.concatenate([[0], np.cumsum(send_sizes)]).astype(np.int64)
        ro = np.concatenate([[0], np.cumsum(recv_sizes)]).astype(np.int64)
        ops = []
        for p in range(self.world):
            sv = send_flat[int(so[p]):int(so[p + 1])]
            rv = recv_flat[int(ro[p]):int(ro[p + 1])]
            if p == self.rank:
                rv.copy_(sv)
                continue
            ops.append(dist.P2POp(dist.isend, sv, self._global_rank(p), self.group))
            ops.append(dist.P2POp(dist.irecv, rv, self._global_rank(p), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _global_rank(self, p):
        return p if self.group is None else self.dist.get_global_rank(self.group, p)

    def _alltoall_blocks(self, send_shapes, recv_shapes, pack):
        """Generic re-slab: pack(p, buf) fills the (flat) block for peer p; returns the received flat blocks."""
        ops = self.ops
        ssz = [int(np.prod(sh)) for sh in send_shapes]
        rsz = [int(np.prod(sh)) for sh in recv_shapes]
        send_flat = ops.empty((sum(ssz),))
        recv_flat = ops.empty((sum(rsz),))
        off = 0
        for p in range(self.world):
            pack(p, send_flat[off:off + ssz[p]])
            off += ssz[p]
        self._all_to_all(send_flat, ssz, recv_flat, rsz)
        out, off = [], 0
        for p in range(self.world):
            out.append(recv_flat[off:off + rsz[p]])
            off += rsz[p]
        return out

    # ------------------------------------------------------------------ Z pass: explicit periodic halo exchange
    def _z_halo_extended(self, slab, r):
        """Returns [r + Zl + r][Y][X]: the local Z-slab with r periodic halo slices on both sides, gathered from the
        owning ranks with grouped send/recv (P2P over NVLink under NCCL)."""
        dist, torch, ops = self.dist, self.torch, self.ops
        Z, Y, X = self.shape
        zs, ze = self.z_range
        Zl = ze - zs
        ext = ops.empty((Zl + 2 * r, Y, X))
        ext[r:r + Zl].copy_(slab)
        if r == 0:
            return ext
        owner_ranges = [split_range(Z, self.world, p) for p in range(self.world)]

        def owner_of(z):
            for p, (a, b) in enumerate(owner_ranges):
                if a <= z < b:
                    return p
            raise AssertionError

        # what I need: ext index e <-> global slice (zs - r + e) mod Z, for e in halo positions
        need = [(e, (zs - r + e) % Z) for e in list(range(r)) + list(range(r + Zl, Zl + 2 * r))]
        # what peer p needs from me, in p's own ext order (both sides enumerate identically -> matched messages)
        p2p = []
        recv_bufs = {}
        for p in range(self.world):
            if p == self.rank:
                continue
            mine = [(e, z) for (e, z) in need if owner_of(z) == p]
            if mine:
                buf = ops.empty((len(mine), Y, X))
                recv_bufs[p] = (buf, mine)
                p2p.append(dist.P2POp(dist.irecv, buf, self._global_rank(p), self.group))
            pzs, pze = owner_ranges[p]
            pl = pze - pzs
            theirs = [(pzs - r + e) % Z for e in list(range(r)) + list(range(r + pl, pl + 2 * r))]
            theirs = [z for z in theirs if zs <= z < ze]
            if theirs:
                sbuf = torch.stack([slab[z - zs] for z in theirs]) if len(theirs) > 1 else slab[theirs[0] - zs][None].contiguous()
                p2p.append(dist.P2POp(dist.isend, sbuf.contiguous(), self._global_rank(p), self.group))
        for (e, z) in need:                      # halo slices that wrap onto my own slab (few ranks / tiny Z)
            if zs <= z < ze:
                ext[e].copy_(slab[z - zs])
        if p2p:
            for w in dist.batch_isend_irecv(p2p):
                w.wait()
        for p, (buf, mine) in recv_bufs.items():
            for i, (e, _z) in enumerate(mine):
                ext[e].copy_(buf[i])
        return ext

    # ------------------------------------------------------------------ the three passes
    def filter(self, slab, kernels, want_zy: bool = False):
        """slab: this rank's Z-slab [Zl][Y][X] (float32, device). Returns (zy_slab or None, zyx_slab), both Z-slabs."""
        ops, G = self.ops, self.world
        Z, Y, X = self.shape
        zs, ze = self.z_range
        ys, ye = self.y_range
        xs, xe = self.x_range
        Zl, Yl, Xl = ze - zs, ye - ys, xe - xs
        kz, ky, kx = (np.asarray(k, np.float64) for k in kernels)
        rz, ry, rx = kz.size // 2, ky.size // 2, kx.size // 2
        zr = [split_range(Z, G, p) for p in range(G)]
        yr = [split_range(Y, G, p) for p in range(G)]
        xr = [split_range(X, G, p) for p in range(G)]
        if tuple(slab.shape) != (Zl, Y, X):
            raise ValueError(f"rank {self.rank} expects a Z-slab of shape {(Zl, Y, X)}, got {tuple(slab.shape)}")

        # ---- Z pass
        ext = self._z_halo_extended(slab, rz)
        A = ops.empty((Zl, Y, X))
        v = View(Zl + 2 * rz, Zl, rz, 0, Y, X, Y * X, X, Y * X, X)
        ops.filter_view(ext, A, v, kz, self.flow, self.chunk, self.exact)
        del ext

        # ---- re-slab Z -> Y (halo included)
        def pack_zy(p, buf):
            a, b = yr[p]
            n = (b - a) + 2 * ry
            ops.copy3d(A, 0, Y * X, X, a - ry, Y, 0, X, buf, 0, n * X, X, Zl, n, X)
        blocks = self._alltoall_blocks([(Zl, (yr[p][1] - yr[p][0]) + 2 * ry, X) for p in range(G)],
                                       [(zr[p][1] - zr[p][0], Yl + 2 * ry, X) for p in range(G)], pack_zy)
        del A
        Ye = Yl + 2 * ry
        Ae = ops.empty((Z, Ye, X))
        flatA = Ae.view(-1)
        for p in range(G):
            a, b = zr[p]
            flatA[a * Ye * X:b * Ye * X].copy_(blocks[p])
        del blocks

        # ---- Y pass
        B = ops.empty((Z, Yl, X))
        v = View(Ye, Yl, ry, 0, Z, X, X, Ye * X, X, Yl * X)
        ops.filter_view(Ae, B, v, ky, self.flow, self.chunk, self.exact)
        del Ae

        zy_slab = None
        if want_zy:   # the reference leaves the Z+Y intermediate in `vol` (src/flowdenoising.py:289)
            def pack_b(p, buf):
                a, b = zr[p]
                ops.copy3d(B, a * Yl * X, Yl * X, X, 0, Yl, 0, X, buf, 0, Yl * X, X, b - a, Yl, X)
            blocks = self._alltoall_blocks([(zr[p][1] - zr[p][0], Yl, X) for p in range(G)],
                                           [(Zl, yr[p][1] - yr[p][0], X) for p in range(G)], pack_b)
            zy_slab = ops.empty((Zl, Y, X))
            for p in range(G):
                a, b = yr[p]
                ops.copy3d(blocks[p], 0, (b - a) * X, X, 0, b - a, 0, X, zy_slab, a * X, Y * X, X, Zl, b - a, X)
            del blocks

        # ---- re-slab Y -> X (halo included, transposed while unpacking)
        def pack_yx(p, buf):
            a, b = xr[p]
            n = (b - a) + 2 * rx
            ops.copy3d(B, 0, Yl * X, X, 0, Yl, a - rx, X, buf, 0, Yl * n, n, Z, Yl, n)
        Xe = Xl + 2 * rx
        blocks = self._alltoall_blocks([(Z, Yl, (xr[p][1] - xr[p][0]) + 2 * rx) for p in range(G)],
                                       [(Z, yr[p][1] - yr[p][0], Xe) for p in range(G)], pack_yx)
        del B
        Ce = ops.empty((Z, Xe, Y))
        for p in range(G):
            a, b = yr[p]
            # block[z][y][x] -> Ce[z][x][a + y]
            ops.transpose_strided(blocks[p], 0, (b - a) * Xe, Xe, Ce, a, Xe * Y, Y, Z, b - a, Xe)
        del blocks

        # ---- X pass
        D = ops.empty((Z, Xl, Y))
        v = View(Xe, Xl, rx, 0, Z, Y, Y, Xe * Y, Y, Xl * Y)
        ops.filter_view(Ce, D, v, kx, self.flow, self.chunk, self.exact)
        del Ce

        # ---- re-slab X -> Z (blocks are contiguous on the sending side; transposed while unpacking)
        def pack_xz(p, buf):
            a, b = zr[p]
            ops.copy3d(D, a * Xl * Y, Xl * Y, Y, 0, Xl, 0, Y, buf, 0, Xl * Y, Y, b - a, Xl, Y)
        blocks = self._alltoall_blocks([(zr[p][1] - zr[p][0], Xl, Y) for p in range(G)],
                                       [(Zl, xr[p][1] - xr[p][0], Y) for p in range(G)], pack_xz)
        del D
        out = ops.empty((Zl, Y, X))
        for p in range(G):
            a, b = xr[p]
            # block[z][x][y] -> out[z][y][a + x]
            ops.transpose_strided(blocks[p], 0, (b - a) * Y, Y, out, a, Y * X, X, Zl, b - a, Y)
        return zy_slab, out
