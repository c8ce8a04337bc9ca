"""Sequence-parallel exchanges over NVLink peer memory — host side (kernels: `csrc/peer.cu`, push epilogues in
`csrc/gemm_tcgen05.cu` / `csrc/fa_tcgen05.cu`).  No reference counterpart (the reference is single-GPU);
the oracle is the single-GPU result.

torch.distributed._symmetric_memory is used for what it is — plumbing: it allocates the exchange buffers and maps every
rank's buffer into every other rank's address space.  The barrier, the strided copy kernel and the push epilogues are libbya.so.

`PeerGroup`      symmetric buffers by name (local tensor + one tensor view per peer), the epoch barrier, segment pushes
`*_segments()`   pure functions building the strided-segment tables of each exchange (unit-tested on CPU against the
                 torch statements of the same permutations in `sp.py`)
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

SEG_DTYPE = np.dtype([("src_off", "<i8"), ("dst_off", "<i8"), ("src_outer_stride", "<i8"), ("dst_outer_stride", "<i8"),
                      ("src_row_stride", "<i8"), ("dst_row_stride", "<i8"), ("peer", "<i4"), ("outer", "<i4"),
                      ("rows", "<i4"), ("row_bytes", "<i4")])   # == ByaPullSeg (include/bya.h), 64 bytes
assert SEG_DTYPE.itemsize == 64


def _seg(peer, src_off, dst_off, outer, rows, row_bytes, src_outer_stride, dst_outer_stride, src_row_stride, dst_row_stride):
    return (src_off, dst_off, src_outer_stride, dst_outer_stride, src_row_stride, dst_row_stride, peer, outer, rows, row_bytes)


def router_gather_segments(P: int, rank: int, CF: int, hwl: int, M: int, Ws: int, esz: int = 2) -> np.ndarray:
    """Every peer s holds s_send [dest][M = CF*hwl][Ws]; this rank wants s_full [(c,f)][P*hwl (all positions)][Ws] for its
    heads: block [(c,f)][hwl] of peer s's slot `rank` lands at positions [s*hwl, (s+1)*hwl) of every (c,f).
    == all_to_all_single + `sp.router_gather_positions` + copy."""
    row = Ws * esz
    return np.array([_seg(s, rank * M * row, s * hwl * row, CF, hwl, row, hwl * row, P * hwl * row, row, row) for s in range(P)],
                    dtype=SEG_DTYPE)


def router_scatter_segments(P: int, rank: int, CF: int, hwl: int, M: int, Wo: int, esz: int = 2) -> np.ndarray:
    """Every peer s holds s_att [(c,f)][P*hwl][Wo] (all positions, ITS heads); this rank wants, for its own positions,
    o_recv [src s][(c,f), local position][Wo] (the K-blocked A operand of the out-projection).
    == `sp.router_scatter_positions` + copy + all_to_all_single."""
    row = Wo * esz
    return np.array([_seg(s, rank * hwl * row, s * M * row, CF, hwl, row, P * hwl * row, hwl * row, row, row) for s in range(P)],
                    dtype=SEG_DTYPE)


def face_query_segments(P: int, rank: int, frames: int, hw: int, text_len: int, rows_per_rank: int, width: int,
                        esz: int = 2) -> np.ndarray:
    """Every rank holds the face queries of ITS token rows in a [rows_per_rank, width] buffer (local row = global row -
    owner * rows_per_rank; global row = text_len + video token).  This rank's router shard needs, for every frame, the
    positions [rank*hwl, (rank+1)*hwl) (clamped to hw-1: padding rows repeat the last position) -> [frames*hwl, width].
    == all_gather of all queries + index_select(`sp.router_local_tokens`)."""
    hwl = (hw + P - 1) // P
    row = width * esz
    segs = []
    for f in range(frames):
        pos = np.minimum(np.arange(hwl) + rank * hwl, hw - 1)
        g = text_len + f * hw + pos                      # global rows wanted, in local order
        j = 0
        while j < hwl:
            owner = int(g[j] // rows_per_rank)
            k = j + 1                                     # extend the run while rows stay consecutive and on one owner
            while k < hwl and g[k] == g[k - 1] + 1 and g[k] // rows_per_rank == owner:
                k += 1
            segs.append(_seg(owner, int(g[j] - owner * rows_per_rank) * row, (f * hwl + j) * row, 1, k - j, row, 0, 0, row, row))
            j = k
    return np.array(segs, dtype=SEG_DTYPE)


def routing_gather_segments(P: int, frames: int, hw: int, chars: int, esz: int = 4) -> np.ndarray:
    """Every peer s holds r_loc [frames*hwl, chars] (its positions); everyone wants routing [frames*hw, chars]."""
    hwl = (hw + P - 1) // P
    row = chars * esz
    segs = []
    for s in range(P):
        n = min(hwl, hw - s * hwl)
        if n > 0:
            segs.append(_seg(s, 0, s * hwl * row, frames, n, row, hwl * row, hw * row, row, row))
    return np.array(segs, dtype=SEG_DTYPE)


def push_table(make_pull, P: int, rank: int) -> np.ndarray:
    """The PUSH table of `rank` from the pull tables of all ranks: whatever destination d would pull from `rank` is what
    `rank` writes into d's buffer (same offsets and strides; `peer` becomes the destination)."""
    out = []
    for d in range(P):
        t = make_pull(d)
        t = t[t["peer"] == rank].copy()
        t["peer"] = d
        out.append(t)
    return np.concatenate(out)


def simulate_push(tables: List[np.ndarray], local_bufs: List[np.ndarray], dst_bufs: List[np.ndarray]):
    """Host reference of `bya_peer_copy(push=1)` run by every rank (tests)."""
    for r, segs in enumerate(tables):
        src = local_bufs[r].view(np.uint8).reshape(-1)
        for s in segs:
            d = dst_bufs[int(s["peer"])].view(np.uint8).reshape(-1)
            for o in range(int(s["outer"])):
                for q in range(int(s["rows"])):
                    a = int(s["src_off"] + o * s["src_outer_stride"] + q * s["src_row_stride"])
                    b = int(s["dst_off"] + o * s["dst_outer_stride"] + q * s["dst_row_stride"])
                    d[b:b + int(s["row_bytes"])] = src[a:a + int(s["row_bytes"])]
    return dst_bufs


def simulate_pull(segs: np.ndarray, peer_bufs: List[np.ndarray], dst: np.ndarray) -> np.ndarray:
    """Host reference of `bya_peer_pull` on byte arrays (tests)."""
    d = dst.view(np.uint8).reshape(-1)
    for s in segs:
        src = peer_bufs[int(s["peer"])].view(np.uint8).reshape(-1)
        for o in range(int(s["outer"])):
            for r in range(int(s["rows"])):
                a = int(s["src_off"] + o * s["src_outer_stride"] + r * s["src_row_stride"])
                b = int(s["dst_off"] + o * s["dst_outer_stride"] + r * s["dst_row_stride"])
                d[b:b + int(s["row_bytes"])] = src[a:a + int(s["row_bytes"])]
    return dst


def vec_bytes_for(segs: np.ndarray) -> int:
    """Widest access (16 / 8 / 4 bytes) every offset, stride and row length of the table allows."""
    vals = []
    for k in ("src_off", "dst_off", "src_outer_stride", "dst_outer_stride", "src_row_stride", "dst_row_stride", "row_bytes"):
        vals += [int(v) for v in segs[k]]
    for v in (16, 8, 4):
        if all(x % v == 0 for x in vals):
            return v
    raise RuntimeError("bya_b200.peer: segment table is not 4-byte aligned")


class PeerGroup:
    """Symmetric exchange buffers of one sequence-parallel group + the device-side epoch barrier."""

    def __init__(self, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.symm_mem = symm_mem
        self.group, self.device = group, device
        self.rank, self.P = dist.get_rank(group), dist.get_world_size(group)
        if self.P > 8:
            raise RuntimeError("bya_b200.peer: at most 8 ranks (one NVSwitch domain)")
        self.bufs: Dict[str, tuple] = {}
        self.tables: Dict[str, tuple] = {}
        flags, peers, self.flag_ptrs = self._alloc((64,), torch.int32)
        flags.zero_()
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)           # every rank's flags are zero before anyone signals
        self.flags = flags
        self.barriers = 0

    def _alloc(self, shape, dtype):
        import torch.distributed as dist

        from .engine import poison_scratch

        t = self.symm_mem.empty(*shape, dtype=dtype, device=self.device)
        hdl = self.symm_mem.rendezvous(t, self.group)
        if poison_scratch() and dtype.is_floating_point:   # debugging aid (engine.poison_scratch): NaN until somebody writes
            t.fill_(float("nan"))
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)                # nobody pushes into a buffer that is still being filled
        peers = [t if r == self.rank else hdl.get_buffer(r, tuple(shape), dtype) for r in range(self.P)]
        ptrs = torch.tensor([p.data_ptr() for p in peers], dtype=torch.int64, device=self.device)
        return t, peers, ptrs

    def get(self, name: str, shape, dtype=torch.bfloat16):
        """Symmetric buffer `name` (collective on first use / shape change: every rank must ask in the same order).
        Returns (local tensor, [tensor view of the same buffer on each rank], device table of the base pointers)."""
        b = self.bufs.get(name)
        if b is None or tuple(b[0].shape) != tuple(shape) or b[0].dtype != dtype:
            b = self._alloc(tuple(shape), dtype)
            self.bufs[name] = b
        return b

    def barrier(self):
        from . import ops

        ops.peer_barrier(self.counter, self.flag_ptrs, self.rank, self.P)
        self.barriers += 1

    def push(self, key, make_pull, local_src: torch.Tensor, dst_ptrs: torch.Tensor, blocks: int = 592):
        """Writes this rank's part of an exchange into the peers' (symmetric) destination buffers.  `make_pull(d)` is the
        strided-segment table rank d would use to PULL the exchange; the push table is derived from it once per `key`.
        `blocks` ~ CTAs in total (posted writes need far fewer bytes in flight than remote reads)."""
        from . import ops

        t = self.tables.get(key)
        if t is None:
            segs = push_table(make_pull, self.P, self.rank)
            dev = torch.from_numpy(segs.view(np.uint8).reshape(-1).copy()).to(self.device)
            t = (dev, len(segs), vec_bytes_for(segs), max(1, min(64, blocks // max(len(segs), 1))))
            self.tables[key] = t
        if t[1]:
            ops.peer_copy(t[0], t[1], dst_ptrs, local_src, True, t[2], t[3])
