"""``DistributedDataParallel``-compatible wrapper whose gradient exchange is a peer-HBM kernel, not NCCL.

Reference surface: ``torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])``
(multi-gpu-distributed-cls.py:341): module pass-through, ``module.``-prefixed ``state_dict`` keys
(:192, :362, test.py:96-101), rank-0 parameter broadcast at wrap time (SP/torch/nn/parallel/distributed.py:879-889),
gradient mean over ranks during ``loss.backward()`` (Reducer, distributed.py:1255-1280).

Design: every rank owns a contiguous 1/world slice of every bucket (embeddings | layer i | head).
Gradients (bf16) and shadow weights (bf16) live in cudaMalloc'ed buffers that every peer maps through CUDA IPC; the
exchange is ``b2_bucket_reduce_adamw``: read my slice from all peers over NVSwitch, mean in fp32, HF-AdamW on my
fp32 master slice, store the new bf16 weights into every peer.  torch.distributed is used only for the one-time
handle exchange / initial broadcast and for re-assembling fp32 masters when a checkpoint is written.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib as L


class _DevBuf:
    """cudaMalloc'ed, IPC-exportable buffer exposed to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, nbytes):
        import ctypes
        p = ctypes.c_void_p()
        L.call("b2_comm_alloc", nbytes, ctypes.byref(p))
        self.ptr, self.nbytes = p.value, nbytes

    def handle(self):
        import ctypes
        buf = ctypes.create_string_buffer(L.IPC_HANDLE_BYTES)
        L.call("b2_comm_export", self.ptr, buf)
        return bytes(buf.raw)

    def tensor(self, dtype, device):
        class _Iface:
            pass
        o = _Iface()
        o.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False),
                                      "version": 2, "strides": None}
        t = torch.as_tensor(o, device=device)
        t._b2_keepalive = self
        return t.view(dtype)


def _import_handle(handle_bytes):
    import ctypes
    p = ctypes.c_void_p()
    L.call("b2_comm_import", handle_bytes, ctypes.byref(p))
    return p.value


class PeerComm:
    """Symmetric buffers of one process group: one entry per name, local pointer + every peer's mapped pointer."""

    def __init__(self, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.local = {}
        self.peers = {}
        self.epochs = torch.zeros(L.FLAG_SLOTS, dtype=torch.int32, device=device)
        self.alloc("flags", L.FLAG_SLOTS * self.world * 4)
        self.alloc("scalar", 2 * self.world * 4)

    def alloc(self, name, nbytes):
        buf = _DevBuf(nbytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, buf.handle(), group=self.group)
        ptrs = []
        for r, h in enumerate(handles):
            ptrs.append(buf.ptr if r == self.rank else _import_handle(h))
        self.local[name] = buf
        self.peers[name] = ptrs
        return buf

    def epoch_ptr(self, slot):
        return self.epochs.data_ptr() + 4 * slot

    def barrier(self, slot, stream):
        L.call("b2_peer_barrier", L.ptr_array(self.peers["flags"]), self.world, self.rank, slot,
               self.epoch_ptr(slot), stream)


# flag slots
_SLOT_GRADS_READY, _SLOT_UPDATE_DONE, _SLOT_LOSS, _SLOT_GATHER, _SLOT_BUCKET0 = 0, 1, 2, 3, 8


class DistributedDataParallel(nn.Module):
    def __init__(self, module, device_ids=None, output_device=None, process_group=None, overlap=True, **unused):
        super().__init__()
        if not hasattr(module, "_engine"):
            raise TypeError("this DistributedDataParallel wraps the b200 BertForSequenceClassification")
        if module._engine is None:
            raise RuntimeError("call model.cuda() before wrapping (as the reference does, :340-341)")
        self.module = module
        self.group = process_group
        if dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        else:
            self.world, self.rank = 1, 0
        self.overlap = overlap
        self.dma = False
        self.comm = None
        self._master_stale = False
        self._side = None
        self._pending = None
        # plain attribute, NOT a registered submodule (model -> wrapper -> model would be a module cycle)
        object.__setattr__(module, "_ddp", self)
        eng = module._engine
        if self.world > 1:
            if self.world > 8:
                raise ValueError("peer-HBM exchange covers one NVSwitch domain (world <= 8)")
            # DDP init sync: parameters of rank 0 win
            dist.broadcast(module._flat, src=0, group=process_group)
            self.comm = PeerComm(eng.dev, process_group)
            n = module._layout.total
            shadow = self.comm.alloc("shadow", 2 * n).tensor(torch.bfloat16, eng.dev)
            grads = self.comm.alloc("grads", 2 * n).tensor(torch.bfloat16, eng.dev)
            eng.rebind(shadow, grads)
            eng.refresh_shadow()
            self._slices = self._make_slices()
            # staging for the DMA form of the exchange: my slice of every bucket as held by each peer
            import os
            self.dma = os.environ.get("B2_DDP_DMA", "1") != "0"
            self._stage_off, off = [], 0
            for (sb, se) in self._slices:
                row = {}
                for r in range(self.world):
                    if r != self.rank:
                        row[r] = off
                        off += (2 * (se - sb) + 255) // 256 * 256
                self._stage_off.append(row)
            self._stage = torch.empty(max(off, 256), dtype=torch.uint8, device=eng.dev)
            self._side = torch.cuda.Stream(device=eng.dev)
            torch.cuda.synchronize(eng.dev)
            dist.barrier(group=process_group)

    def _make_slices(self):
        out = []
        for (b, e, _label) in self.module._layout.buckets:
            per = ((e - b) // 8 + self.world - 1) // self.world * 8
            sb = min(e, b + self.rank * per)
            se = min(e, sb + per)
            out.append((sb, se))
        return out

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ---- hooks called by the engine during backward (autograd thread) -----------------------------------------------
    def _bucket_ready(self, idx):
        """Bucket `idx` holds this rank's final local gradients.  With an optimizer attached and overlap on, start
        its exchange + update on the side stream right away so it hides behind the rest of backward."""
        opt = self.module._optimizer
        if self.world == 1 or not self.overlap or opt is None or not getattr(opt, "_armed", False):
            return
        eng = self.module._engine
        main = torch.cuda.current_stream(eng.dev)
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        s = self._side.cuda_stream
        self.comm.barrier(_SLOT_BUCKET0 + idx, s)
        self._exchange_update(opt, idx, s)
        if self._pending is None:
            self._pending = set()
        self._pending.add(idx)

    def _exchange_update(self, opt, idx, s):
        """Mean over ranks + HF-AdamW on my slice of bucket `idx` + delivery of the new bf16 weights to every rank.
        Kernel form: the reduce kernel loads the peers' slices / stores the peers' shadows itself through the mapped
        pointers.  DMA form (default for every bucket but the last one produced): the transfers are copy-engine
        copies over NVLink -- they run beside the GEMM CTAs instead of time-slicing with them (a 640-thread GEMM CTA
        owns its SM's registers, so an SM-driven exchange kernel can only run between them) -- and the reduce kernel
        works on local memory only."""
        sb, se = self._slices[idx]
        if se <= sb:
            return
        peers_g, peers_s = self.comm.peers["grads"], self.comm.peers["shadow"]
        if not self.dma or idx == 0:
            opt.update_range(sb, se, self.world, self.rank, peers_g, peers_s, s)
            return
        nbytes = 2 * (se - sb)
        g_ptrs, s_ptrs = [], []
        for r in range(self.world):
            if r == self.rank:
                g_ptrs.append(peers_g[r])
                s_ptrs.append(peers_s[r])
                continue
            st = self._stage.data_ptr() + self._stage_off[idx][r]
            L.call("b2_copy_async", st, peers_g[r] + 2 * sb, nbytes, s)
            g_ptrs.append(st - 2 * sb)        # the kernel indexes base + absolute element index
            s_ptrs.append(None)
        opt.update_range(sb, se, self.world, self.rank, g_ptrs, s_ptrs, s)
        mine = peers_s[self.rank] + 2 * sb
        for r in range(self.world):
            if r != self.rank:
                L.call("b2_copy_async", peers_s[r] + 2 * sb, mine, nbytes, s)

    def _on_backward_done(self):
        pass

    def _optimizer_step(self, opt):
        """world > 1 body of ``optimizer.step()``."""
        eng = self.module._engine
        main = torch.cuda.current_stream(eng.dev)
        nb = len(self.module._layout.buckets)
        done = self._pending or set()
        if len(done) == nb:
            # everything was launched from the backward hooks: just join
            s = self._side.cuda_stream
            self.comm.barrier(_SLOT_UPDATE_DONE, s)
            opt.advance(s)
            ev = torch.cuda.Event()
            ev.record(self._side)
            main.wait_event(ev)
        else:
            if done:
                ev = torch.cuda.Event()
                ev.record(self._side)
                main.wait_event(ev)
            s = main.cuda_stream
            self.comm.barrier(_SLOT_GRADS_READY, s)
            for idx in range(nb):
                if idx in done:
                    continue
                self._exchange_update(opt, idx, s)
            self.comm.barrier(_SLOT_UPDATE_DONE, s)
            opt.advance(s)
        self._pending = None
        self._master_stale = True
        opt._armed = bool(self.overlap) and not getattr(opt, "_amp_seen", False)

    def _gather_master(self):
        """fp32 masters are updated slice-wise by their owner ranks; re-assemble them (checkpoint time only)."""
        if self.world == 1 or not self._master_stale:
            return
        flat = self.module._flat
        for (b, e, _label) in self.module._layout.buckets:
            per = ((e - b) // 8 + self.world - 1) // self.world * 8
            for r in range(self.world):
                sb = min(e, b + r * per)
                se = min(e, sb + per)
                if se > sb:
                    dist.broadcast(flat[sb:se], src=r, group=self.group)
        self._master_stale = False

    # ---- the two small collectives of the reference Trainer -------------------------------------------------------------
    def loss_reduce(self, loss):
        """mean over ranks of a scalar loss (Trainer.loss_reduce, :139-143)."""
        if self.world == 1:
            return loss.clone()
        eng = self.module._engine
        src = loss.detach().to(torch.float32).reshape(1).contiguous()
        dst = torch.empty(1, dtype=torch.float32, device=eng.dev)
        L.call("b2_scalar_allreduce_mean", src.data_ptr(), dst.data_ptr(), L.ptr_array(self.comm.peers["scalar"]),
               L.ptr_array(self.comm.peers["flags"]), self.world, self.rank, _SLOT_LOSS,
               self.comm.epoch_ptr(_SLOT_LOSS), eng.stream())
        return dst.reshape(())

    def all_gather_rows(self, t):
        """rank-ordered concatenation along dim 0 (Trainer.output_reduce, :145-155)."""
        if self.world == 1:
            return t.clone()
        eng = self.module._engine
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        if nbytes % 4 != 0:
            raise ValueError("all_gather_rows: payload must be a multiple of 4 bytes")
        # two buffers per payload size, alternated per call: a fast rank's next store can never land in a buffer a
        # slow rank is still reading (to reach call n+2 it must pass barrier n+1, which the slow rank only joins
        # after its read of call n was enqueued ahead of it on the same stream)
        self._gather_calls = getattr(self, "_gather_calls", 0) + 1
        key = "gather_%d_%d" % (nbytes, self._gather_calls & 1)
        if key not in self.comm.local:
            self.comm.alloc(key, nbytes * self.world)
        L.call("b2_allgather_rows", t.data_ptr(), nbytes, L.ptr_array(self.comm.peers[key]),
               L.ptr_array(self.comm.peers["flags"]), self.world, self.rank, _SLOT_GATHER,
               self.comm.epoch_ptr(_SLOT_GATHER), eng.stream())
        full = self.comm.local[key].tensor(t.dtype, eng.dev)
        return full.view((self.world * t.shape[0],) + tuple(t.shape[1:])).clone()
