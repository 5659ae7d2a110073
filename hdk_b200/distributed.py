"""Multi-GPU layer of the hot path: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch; gloo on CPU for the host-logic tests).  HDK itself has none of this — fragments are all
stamped with device 0 (omniscidb/ArrowStorage/ArrowStorage.cpp:365-367) and multi-device results
are merged on the host (QE/Execute.cpp:1224-1336).  Three exchange patterns:

  * perfect-hash group-by  : every rank scans its fragments into a NEUTRAL work table
                             (hdk_b200_launch_partial) → all-reduce per merge class
                             (int64 SUM | fp64 SUM | MIN | MAX) → hdk_b200_finalize
  * baseline-hash group-by : rows are re-partitioned by floor(hi32(MurmurHash64A(key)) * world / 2^32)
                             (hdk_b200_shuffle_count / _scatter, model: QE/RelAlgExecutor.cpp:691-838)
                             → all-to-all → local aggregate; results are disjoint, no merge
  * small join build side  : built once on rank 0, broadcast (table + inner columns)

`PeerExchange` replaces the all-reduce of the first pattern by the library's own exchange over peer memory
(hdk_b200_launch_exchange: the scan kernel's last CTA stores this GPU's partial table into every peer's exchange
buffer through NVLink and raises a flag; every finalize kernel waits for the flags and merges) — NCCL then only
carries the one-time IPC handle exchange.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def world() -> int:
    return dist.get_world_size() if is_dist() else 1


def rank() -> int:
    return dist.get_rank() if is_dist() else 0


def shard_fragments(n_fragments: int, r: Optional[int] = None, w: Optional[int] = None) -> List[int]:
    """fragment i → rank i mod world (what upstream intended with Fragment::deviceIds)."""
    r = rank() if r is None else r
    w = world() if w is None else w
    return [i for i in range(n_fragments) if i % w == r]


def allreduce_work_table(work: torch.Tensor, n_cells: int, sum_i64_cells: int, sum_cells: int, min_cells: int,
                         max_cells: int, group=None) -> None:
    """In-place merge of a neutral work table (int64 view of the scratch buffer) across ranks.
    Layout (hdk_b200_work_table_layout): [0, sum_i64) int64 SUM | [sum_i64, sum) fp64 SUM |
    next min_cells int64 MIN (fp MIN is stored order-encoded) | last max_cells int64 MAX."""
    if not is_dist() or world() == 1:
        return
    w = work.view(torch.int64)[:n_cells]
    if sum_i64_cells:
        dist.all_reduce(w[:sum_i64_cells], op=dist.ReduceOp.SUM, group=group)
    if sum_cells > sum_i64_cells:
        dist.all_reduce(w[sum_i64_cells:sum_cells].view(torch.float64), op=dist.ReduceOp.SUM, group=group)
    if min_cells:
        dist.all_reduce(w[sum_cells:sum_cells + min_cells], op=dist.ReduceOp.MIN, group=group)
    if max_cells:
        dist.all_reduce(w[sum_cells + min_cells:sum_cells + min_cells + max_cells], op=dist.ReduceOp.MAX, group=group)


def all_to_all_rows(cols: List[torch.Tensor], send_counts: torch.Tensor, widths: List[int], group=None):
    """Exchange partition-contiguous column buffers (output of hdk_b200_shuffle_scatter).
    cols[c] is a uint8 tensor holding rows grouped by destination rank; send_counts[r] rows go to rank r.
    Returns (received columns, rows received)."""
    w = world()
    send = send_counts.to(torch.int64)
    recv = torch.empty_like(send)
    if w > 1:
        gathered = [torch.empty_like(send) for _ in range(w)]
        dist.all_gather(gathered, send, group=group)     # counts matrix: recv[r] = send_of_rank_r[me]
        recv = torch.stack([g[rank()] for g in gathered])
    else:
        recv.copy_(send)
    send_l, recv_l = send.tolist(), recv.tolist()
    n_recv = int(sum(recv_l))
    out = []
    for c, width in zip(cols, widths):
        dst = torch.empty(max(n_recv, 1) * width, dtype=torch.uint8, device=c.device)
        if w > 1:
            _all_to_all_bytes(dst[: n_recv * width], c[: int(sum(send_l)) * width], [x * width for x in recv_l],
                              [x * width for x in send_l], group)
        else:
            dst[: n_recv * width].copy_(c[: n_recv * width])
        out.append(dst)
    return out, n_recv


def _all_to_all_bytes(dst, src, recv_sizes, send_sizes, group=None):
    """all_to_all_single over NCCL; gloo has no all-to-all, so the CPU tests fall back to isend / irecv pairs."""
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(dst, src, output_split_sizes=recv_sizes, input_split_sizes=send_sizes, group=group)
        return
    r, ops, so, ro = rank(), [], 0, 0
    for peer in range(world()):
        s_chunk, r_chunk = src[so:so + send_sizes[peer]], dst[ro:ro + recv_sizes[peer]]
        so += send_sizes[peer]
        ro += recv_sizes[peer]
        if peer == r:
            r_chunk.copy_(s_chunk)
            continue
        if send_sizes[peer]:
            ops.append(dist.isend(s_chunk.contiguous(), peer, group=group))
        if recv_sizes[peer]:
            ops.append(dist.irecv(r_chunk, peer, group=group))
    for o in ops:
        o.wait()


def broadcast_tensor(t: Optional[torch.Tensor], nbytes: int, device, src: int = 0, group=None) -> torch.Tensor:
    """Broadcast a byte buffer (join hash table / dimension column) from `src` to every rank."""
    if rank() != src or t is None:
        t = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if is_dist() and world() > 1:
        dist.broadcast(t, src=src, group=group)
    return t


class PeerExchange:
    """Peer-visible exchange buffers of one plan (hdk_b200_peer_alloc + CUDA IPC handles gathered through the process
    group).  `ptrs()` is the host array hdk_b200_launch_exchange wants; `next_epoch()` the per-launch counter."""

    def __init__(self, lib, plan, qmd, device, group=None):
        import ctypes as C

        from . import _lib
        self.lib, self.rank, self.world = lib, rank(), world()
        nbytes = C.c_size_t(0)
        _lib.check(lib.hdk_b200_exchange_bytes(C.byref(plan), C.byref(qmd), self.world, C.byref(nbytes)), "exchange_bytes")
        self.nbytes = nbytes.value
        self.local = C.c_void_p(0)
        handle = (C.c_uint8 * 64)()
        _lib.check(lib.hdk_b200_peer_alloc(self.nbytes, C.byref(self.local), handle), "peer_alloc")
        _lib.check(lib.hdk_b200_exchange_init(self.local, None), "exchange_init")
        torch.cuda.synchronize(device)
        self.opened = []
        self._ptrs = (C.c_void_p * self.world)()
        self._ptrs[self.rank] = self.local.value
        if self.world > 1:
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            for r, g in enumerate(gathered):
                if r == self.rank:
                    continue
                hb = (C.c_uint8 * 64)(*g.cpu().tolist())
                p = C.c_void_p(0)
                _lib.check(lib.hdk_b200_peer_open(hb, C.byref(p)), f"peer_open(rank {r})")
                self.opened.append(p)
                self._ptrs[r] = p.value
            dist.barrier(group=group)    # every rank's flags are zeroed and mapped before the first launch
        self.epoch = 0

    def ptrs(self):
        return self._ptrs

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def close(self):
        for p in self.opened:
            self.lib.hdk_b200_peer_close(p)
        self.opened = []
        if self.local:
            self.lib.hdk_b200_peer_free(self.local)
            self.local = None


class _DevMem:
    """zero-copy torch view of raw device memory (cudaMalloc through the library)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerRows:
    """Peer-visible receive buffers of a partitioned aggregation, one per plan column, `capacity_rows` rows each.
    hdk_b200_shuffle_scatter_to writes every partition straight into its owner's buffers through NVLink — the
    scatter kernel is the all-to-all; only the per-partition counts travel through the process group."""

    def __init__(self, lib, widths, capacity_rows: int, device, group=None):
        import ctypes as C

        from . import _lib
        self.lib, self.widths, self.capacity, self.device, self.group = lib, list(widths), int(capacity_rows), device, group
        self.rank, self.world = rank(), world()
        n = len(self.widths)
        self.local_ptrs, handles = [], []
        for w in self.widths:
            p, h = C.c_void_p(0), (C.c_uint8 * 64)()
            _lib.check(lib.hdk_b200_peer_alloc(max(self.capacity, 1) * w, C.byref(p), h), "peer_alloc")
            self.local_ptrs.append(p)
            handles.append(bytes(h))
        self.opened = []
        table = [[0] * n for _ in range(self.world)]
        table[self.rank] = [p.value for p in self.local_ptrs]
        if self.world > 1:
            mine = torch.tensor(list(b"".join(handles)), dtype=torch.uint8, device=device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            for r, g in enumerate(gathered):
                if r == self.rank:
                    continue
                raw = bytes(g.cpu().tolist())
                for c in range(n):
                    hb = (C.c_uint8 * 64)(*raw[64 * c: 64 * (c + 1)])
                    p = C.c_void_p(0)
                    _lib.check(lib.hdk_b200_peer_open(hb, C.byref(p)), f"peer_open(rank {r}, column {c})")
                    self.opened.append(p)
                    table[r][c] = p.value
            dist.barrier(group=group)
        self.dest = torch.tensor([x for row in table for x in row], dtype=torch.int64, device=device)   # [world * n_cols]

    def local_column(self, c: int, rows: int):
        return torch.as_tensor(_DevMem(self.local_ptrs[c].value, max(rows, 1) * self.widths[c]), device=self.device)[: rows * self.widths[c]]

    def close(self):
        """Collective: every rank unmaps its peers' buffers, THEN (after a barrier) frees its own — freeing memory a
        slower peer still has IPC-mapped is undefined behaviour (cudaIpcOpenMemHandle contract)."""
        for p in self.opened:
            self.lib.hdk_b200_peer_close(p)
        self.opened = []
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self.local_ptrs:
            self.lib.hdk_b200_peer_free(p)
        self.local_ptrs = []
