"""NVLink peer-memory buffers for the fused TP all-reduce kernel (csrc/sq_tp.cu).

Each rank cudaMalloc's one block [partial A | partial B | reduced A | reduced B | flags | epoch | row flags A | row flags
B], exports it with CUDA IPC, and maps every
peer's block (cudaIpcOpenMemHandle, peer access over NVLink / NVSwitch).  torch sees the two partial buffers as ordinary
fp16 tensors (zero-copy via __cuda_array_interface__), so the row-parallel GEMMs write their outputs straight into
peer-visible memory with `torch.mm(..., out=...)`."""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr, stream_ptr


class _CudaArray:
    def __init__(self, address: int, shape, typestr="<f2"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (address, False), "version": 2}


class PeerBuffers:
    FLAG_BYTES = 1024
    MBOX_WORDS = 8192          # 4-byte payload words per message (tokens + position ids of M <= 2040 plus the state word)

    def __init__(self, group, device, n_max: int, hidden: int):
        lib = _lib.load()
        self.group, self.device = group, torch.device(device)
        self.N, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert 2 <= self.N <= 8
        self.n_max, self.hidden = n_max, hidden
        part = n_max * hidden * 2
        rowflag_bytes = ((n_max * 4 + 1023) // 1024) * 1024
        # one-shot PUSH (small payloads): receive slots for <= push_rows rows from each source, per buffer parity
        self.push_rows = min(n_max, 256)
        recv_bytes = self.N * self.push_rows * hidden * 2
        pflag_bytes = ((self.N * self.push_rows * 4 + 1023) // 1024) * 1024
        # LL two-shot (small payloads): gather area (N sources x rows owned) and reduced-row area, 2 bytes of slot per byte of
        # payload, per buffer parity
        # (N <= 3: one-shot form -- every row from every source, one NVLink trip; else rows owned by rank r % N, two trips)
        self.ll_own = self.push_rows if self.N <= 3 else (self.push_rows + self.N - 1) // self.N
        ll1_bytes = self.N * self.ll_own * hidden * 4
        ll2_bytes = self.push_rows * hidden * 4
        # mailboxes for the driver -> follower messages: 2 channels x 2 parities x MBOX_WORDS 8-byte LL words
        mbox_bytes = 2 * 2 * self.MBOX_WORDS * 8
        total = (4 * part + 2 * self.FLAG_BYTES + 2 * rowflag_bytes + 2 * recv_bytes + pflag_bytes + 2 * (ll1_bytes + ll2_bytes)
                 + mbox_bytes)
        # one-shot (every rank pulls all partials) below 4 ranks, two-shot (reduce-scatter + all-gather in one kernel)
        # from 4 ranks up; SQ_TP_SHOT=1|2 overrides
        shot = os.environ.get("SQ_TP_SHOT", "")
        self.two_shot = (shot == "2") or (shot != "1" and self.N >= 4)
        # shot 3 = one-shot PUSH for small payloads (<= 8 MB pushed per rank).  Opt-in: measured at TP-2 on c2 it is slower
        # than the pull kernel (4.02 vs 3.68 ms / step; 18.5 vs 12.7 us per reduction) -- the system fence between the remote
        # stores and the flag costs the round trip that the pull spends on its loads.
        self.push_ok = shot == "3"
        # shot 4 = LL (data and epoch in one 8-byte store, readers poll).  Measured on c2 (128 rows x 4096, 1 MB payload,
        # profiles/r02_bench_c2_tp*.json): TP-2 one-shot pull 3.69 ms / step vs LL 3.75; TP-4 LL 3.53 vs two-shot pull 3.68;
        # TP-8 two-shot pull 3.40 vs LL 3.69 (LL doubles the bytes, and at 8 ranks each owner gathers 7 x 256 KB).  Defaults
        # follow the measurements: pull below 4 ranks, LL for 4..7 ranks, two-shot pull at 8.
        self.ll_ok = shot == "4" or (shot == "" and 4 <= self.N < 8)
        self.ll_bytes_max = 4 << 20
        self.push_bytes_max = 8 << 20
        base = C.c_void_p()
        check(lib.sq_tp_alloc(C.byref(base), total), "sq_tp_alloc")
        self.base = base.value
        handle = (C.c_uint8 * 64)()
        check(lib.sq_tp_ipc_export(base, handle), "sq_tp_ipc_export")
        handles = [None] * self.N
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.bases = []
        self._opened = []
        for r in range(self.N):
            if r == self.rank:
                self.bases.append(self.base)
                continue
            p = C.c_void_p()
            h = (C.c_uint8 * 64).from_buffer_copy(handles[r])
            check(lib.sq_tp_ipc_open(h, C.byref(p)), "sq_tp_ipc_open")
            self.bases.append(p.value)
            self._opened.append(p.value)
        arr = C.c_void_p * 8
        self.proj_ptrs = [arr(*[(b + w * part) for b in self.bases] + [None] * (8 - self.N)) for w in range(2)]
        self.red_ptrs = [arr(*[(b + (2 + w) * part) for b in self.bases] + [None] * (8 - self.N)) for w in range(2)]
        self.flag_ptrs = arr(*[(b + 4 * part) for b in self.bases] + [None] * (8 - self.N))
        self.epoch_ptr = self.base + 4 * part + self.FLAG_BYTES
        rf0 = 4 * part + 2 * self.FLAG_BYTES
        self.rowflag_ptrs = [arr(*[(b + rf0 + w * rowflag_bytes) for b in self.bases] + [None] * (8 - self.N)) for w in range(2)]
        rv0 = rf0 + 2 * rowflag_bytes
        self.recv_ptrs = [arr(*[(b + rv0 + w * recv_bytes) for b in self.bases] + [None] * (8 - self.N)) for w in range(2)]
        self.pflag_ptrs = arr(*[(b + rv0 + 2 * recv_bytes) for b in self.bases] + [None] * (8 - self.N))
        ll0 = rv0 + 2 * recv_bytes + pflag_bytes
        self.ll1_ptrs = [arr(*[(b + ll0 + w * (ll1_bytes + ll2_bytes)) for b in self.bases] + [None] * (8 - self.N)) for w in range(2)]
        self.ll2_ptrs = [arr(*[(b + ll0 + w * (ll1_bytes + ll2_bytes) + ll1_bytes) for b in self.bases] + [None] * (8 - self.N))
                         for w in range(2)]
        mb0 = ll0 + 2 * (ll1_bytes + ll2_bytes)
        self.mbox_local = [self.base + mb0 + ch * 2 * self.MBOX_WORDS * 8 for ch in range(2)]
        self.mbox_peers = [arr(*([b + mb0 + ch * 2 * self.MBOX_WORDS * 8 for i, b in enumerate(self.bases) if i != self.rank]
                                 + [None] * (8 - (self.N - 1)))) for ch in range(2)]
        self.msg_epoch = torch.zeros(4, dtype=torch.int32, device=self.device)      # [channel] message counters of this rank
        self.msg_on = os.environ.get("SQ_TP_MSG", "ll") == "ll"                      # SQ_TP_MSG=nccl keeps the NCCL broadcasts
        self.buf = [torch.as_tensor(_CudaArray(self.base + w * part, (n_max, hidden)), device=self.device) for w in range(2)]
        assert self.buf[0].data_ptr() == self.base and self.buf[0].dtype == torch.float16
        dist.barrier(group=group)                      # every rank has mapped every peer before the first kernel runs

    def allreduce_add_rmsnorm(self, which: int, resid: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, n: int,
                              eps: float):
        """resid += sum over ranks of partial buffer `which`; out = rmsnorm(resid) * weight  (one kernel per rank)."""
        if self.ll_ok and n <= self.push_rows and n * self.hidden * 2 <= self.ll_bytes_max:
            check(_lib.load().sq_tp_allreduce_ll_add_rmsnorm(ptr(resid), self.buf[which].data_ptr(), self.ll1_ptrs[which],
                                                             self.ll2_ptrs[which], self.epoch_ptr, self.rank, self.N,
                                                             self.push_rows, self.ll_own, ptr(weight), ptr(out), n, self.hidden,
                                                             eps, stream_ptr()), "sq_tp_allreduce_ll_add_rmsnorm")
            return
        if self.push_ok and n <= self.push_rows and (self.N - 1) * n * self.hidden * 2 <= self.push_bytes_max:
            check(_lib.load().sq_tp_allreduce3_add_rmsnorm(ptr(resid), self.buf[which].data_ptr(), self.recv_ptrs[which],
                                                           self.pflag_ptrs, self.epoch_ptr, self.rank, self.N, self.push_rows,
                                                           ptr(weight), ptr(out), n, self.hidden, eps, stream_ptr()),
                  "sq_tp_allreduce3_add_rmsnorm")
            return
        if self.two_shot:
            check(_lib.load().sq_tp_allreduce2_add_rmsnorm(ptr(resid), self.proj_ptrs[which], self.red_ptrs[which],
                                                           self.flag_ptrs, self.rowflag_ptrs[which], self.epoch_ptr,
                                                           self.rank, self.N, ptr(weight), ptr(out), n, self.hidden, eps,
                                                           stream_ptr()), "sq_tp_allreduce2_add_rmsnorm")
            return
        check(_lib.load().sq_tp_allreduce_add_rmsnorm(ptr(resid), self.proj_ptrs[which], self.flag_ptrs, self.epoch_ptr,
                                                      self.rank, self.N, ptr(weight), ptr(out), n, self.hidden, eps,
                                                      stream_ptr()), "sq_tp_allreduce_add_rmsnorm")

    @staticmethod
    def _words(t):
        assert t.is_contiguous() and (t.numel() * t.element_size()) % 4 == 0
        return t.numel() * t.element_size() // 4

    def fits(self, tensors) -> bool:
        """True if the message fits a mailbox (else the caller keeps the NCCL broadcasts; both sides see the same sizes)."""
        return sum(self._words(t) for t in tensors) <= self.MBOX_WORDS

    def publish(self, channel: int, tensors):
        """Rank 0: write `tensors` (<= 3) as LL words into every follower's mailbox of `channel` (one kernel)."""
        ts = list(tensors) + [None] * (3 - len(tensors))
        a = [(ptr(t), self._words(t)) if t is not None else (None, 0) for t in ts]
        check(_lib.load().sq_tp_ll_publish(self.mbox_peers[channel], self.N - 1, self.MBOX_WORDS,
                                           self.msg_epoch.data_ptr() + 4 * channel, a[0][0], a[0][1], a[1][0], a[1][1], a[2][0],
                                           a[2][1], stream_ptr()), "sq_tp_ll_publish")

    def consume(self, channel: int, tensors):
        """Followers: poll the mailbox of `channel` and scatter the words into `tensors` (one kernel)."""
        ts = list(tensors) + [None] * (3 - len(tensors))
        a = [(ptr(t), self._words(t)) if t is not None else (None, 0) for t in ts]
        check(_lib.load().sq_tp_ll_consume(self.mbox_local[channel], self.MBOX_WORDS, self.msg_epoch.data_ptr() + 4 * channel,
                                           self.epoch_ptr + 8, a[0][0], a[0][1], a[1][0], a[1][1], a[2][0], a[2][1], stream_ptr()),
              "sq_tp_ll_consume")

    def error(self) -> int:
        t = torch.as_tensor(_CudaArray(self.epoch_ptr, (4,), "<i4"), device=self.device)
        return int(t[2])
