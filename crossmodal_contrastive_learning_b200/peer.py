"""Row-shard exchange through NVLink peer memory (SURVEY.md section 8 rows e / f2): the `exchange="peer"` option of the
row-sharded criterion.  One process per GPU on ONE node; every rank owns a buffer holding the whole stacked matrix plus the
row statistics, mapped into every other rank with CUDA IPC, and `crossclr_peer_exchange` (csrc/peer.cu) replaces each NCCL
all-gather by one kernel: peer stores of the rank's own slice into all buffers + a cross-rank flag barrier.  torch.distributed
is used once, to hand the 64-byte IPC handles around."""
from __future__ import annotations

import ctypes

import torch

from . import _native as N

_HANDLE = 64
_CTRL_BYTES = 1024          # flags: 2 x 16 words at 0; state: 4 words at 512


class _Raw:
    """CUDA array interface over a raw device allocation (zero-copy torch view)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerPlan:
    """Buffers of one (feature bytes, statistics bytes) geometry on one process group."""

    def __init__(self, group, device, feat_bytes, stats_bytes):
        import torch.distributed as dist
        self.lib = N.load()
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.feat_bytes = (feat_bytes + 255) // 256 * 256
        self.stats_bytes = (stats_bytes + 255) // 256 * 256
        self.generation = 0
        total = self.feat_bytes + self.stats_bytes
        # Every step below is taken by EVERY rank, whatever fails locally: a rank that left the sequence early would leave the
        # others waiting in a collective.  The closing all-reduce is both the agreement on success and the barrier "every rank
        # has mapped every buffer" that must precede the first store.
        err, mine = None, None
        self._local = self._ctrl = None
        self._peers, bases, flags = [], [], []
        with torch.cuda.device(device):
            try:
                self._local = self._alloc(total)
                self._ctrl = self._alloc(_CTRL_BYTES)
                mine = (self._export(self._local), self._export(self._ctrl))
            except Exception as exc:             # noqa: BLE001 -- reported to every rank below
                err = exc
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            if err is None and any(h is None for h in handles):
                err = RuntimeError("a peer rank could not allocate or export its buffers")
            if err is None:
                try:
                    for r, (hb, hc) in enumerate(handles):
                        if r == self.rank:
                            bases.append(self._local); flags.append(self._ctrl)
                        else:
                            pb, pc = self._import(hb), self._import(hc)
                            self._peers += [pb, pc]
                            bases.append(pb); flags.append(pc)
                except Exception as exc:         # noqa: BLE001
                    err = exc
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                raise RuntimeError(f"exchange='peer': the peer buffers could not be mapped on every rank "
                                   f"({'this rank: ' + str(err) if err is not None else 'another rank failed'})")
        self._bases = (ctypes.c_void_p * self.world)(*bases)
        self._flags = (ctypes.c_void_p * self.world)(*flags)
        self._state = self._ctrl + 512
        self._keep = _Raw(self._local, total)
        with torch.cuda.device(device):
            self.bytes_view = torch.as_tensor(self._keep, device=device)

    def _alloc(self, nbytes):
        p = ctypes.c_void_p()
        N.check(self.lib.crossclr_peer_alloc(nbytes, ctypes.byref(p)), "crossclr_peer_alloc")
        return p.value

    def _export(self, ptr):
        buf = ctypes.create_string_buffer(_HANDLE)
        N.check(self.lib.crossclr_peer_export(ptr, buf), "crossclr_peer_export")
        return bytes(buf.raw)

    def _import(self, handle):
        p = ctypes.c_void_p()
        N.check(self.lib.crossclr_peer_import(handle, ctypes.byref(p)), "crossclr_peer_import")
        return p.value

    def feat(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        return self.bytes_view[:n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)

    def stats(self, rows):
        return self.bytes_view[self.feat_bytes:self.feat_bytes + rows * 8].view(torch.float32).view(rows, 2)

    def exchange(self, offset, nbytes, entry_barrier, stream):
        N.check(self.lib.crossclr_peer_exchange(self._bases, self._flags, self.world, self.rank, offset, nbytes,
                                                1 if entry_barrier else 0, self._state, stream), "crossclr_peer_exchange")

    def wire_bytes(self, nbytes):
        """bytes this rank stores over NVLink for one exchange of an `nbytes` slice"""
        return nbytes * (self.world - 1)


_plans = {}


def plan_for(group, device, feat_bytes, stats_bytes):
    """The cached plan of this geometry (creation is collective: every rank of the group must get here together, outside
    CUDA-graph capture -- the first eager step does it)."""
    key = (id(group), str(device), feat_bytes, stats_bytes)
    p = _plans.get(key)
    if p is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("exchange='peer': run one eager step before capturing a CUDA graph (the peer buffers are "
                               "mapped on first use)")
        p = _plans[key] = PeerPlan(group, device, feat_bytes, stats_bytes)
    return p
