"""Single-trial data-parallel training (BASELINE.json configs[3]: one trial, per-rank batch, NCCL gradient all-reduce).

Every rank holds the same weights and AdamW state and its own shard of the training rows.  Each loss phase of each batch
is one split-phase kernel launch that exports the phase's gradient vector instead of applying it
(`raae_train_phase`), one `torch.distributed.all_reduce` (mean) of that vector — 5 all-reduces of 118–238 KB per step,
latency-bound — and one fused AdamW launch (`raae_apply_adam`).  BatchNorm batch statistics and Kendall
pairs stay rank-local (DistributedDataParallel semantics, not large-batch semantics); the BatchNorm running buffers
are averaged across ranks before every validation block so that all ranks score the same model.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .engine import Engine
from .trainer import build_modules


def shard_rows(n_rows, world, rank):
    """Contiguous, equally sized shards (the remainder rows are dropped so that every rank runs the same number of
    batches — a collective per phase needs that)."""
    per = n_rows // world
    return rank * per, (rank + 1) * per


class DataParallelTrainer:
    def __init__(self, cfg, spec_train, aux_train, spec_val, aux_val, device, rank=0, world=1, seed=0):
        self.rank, self.world = rank, world
        self.cfg = dict(cfg)
        self.cfg.setdefault("epoch_stop_smooth", 500)
        lo, hi = shard_rows(len(spec_train), world, rank)
        self.engine = Engine(self.cfg, n_trials=1, device=device, max_rows=max(int(cfg["batch_size"]), len(spec_val)),
                             seeds=[seed * 1000 + rank])            # rank-local noise / dropout streams
        self.modules = build_modules(self.cfg, seed=seed)             # identical initial weights on every rank
        self.engine.load_modules(0, *self.modules)
        self.engine.bind_dataset(spec_train[lo:hi], aux_train[lo:hi], spec_val, aux_val)
        lay = self.engine.lay
        self.grads = [torch.zeros(1, lay.opt[o].n, dtype=torch.float32, device=self.engine.device) for o in range(L.NUM_PHASES)]
        self._gptr = (L._p * L.NUM_PHASES)()
        bs = int(cfg["batch_size"])
        self.n_steps = (self.engine.n_train + bs - 1) // bs

    def _allreduce(self, t):
        if self.world > 1:
            torch.distributed.all_reduce(t)
            t /= self.world

    def _sync_bn_buffers(self):
        if self.world == 1:
            return
        lay, st = self.engine.lay, self.engine.state[0]
        for ni in (0, 1):
            n = lay.net[ni]
            nbn = n.n_linear if ni == 0 else n.n_linear - 1
            lo, hi = n.rm_off[0], n.rv_off[nbn - 1] + n.out_dim[nbn - 1]
            self._allreduce(st[lo:hi])

    def train_epoch(self, epoch, perm=None):
        """One epoch: every batch runs its five phases as launch -> all-reduce -> AdamW.  Returns (losses[12], metrics[6])."""
        eng = self.engine
        if perm is None:
            perm = eng.make_perm(1)[0]
        stop_smooth = float(self.cfg["epoch_stop_smooth"])
        for s in range(self.n_steps):
            for o in range(L.NUM_PHASES):
                if o == 4 and epoch >= stop_smooth:
                    continue
                for k in range(L.NUM_PHASES):
                    self._gptr[k] = self.grads[k].data_ptr() if k == o else None
                L.check(eng.lib.raae_train_phase(eng.handle, int(epoch), s, 1 << o, perm.data_ptr(), self._gptr, eng.stream))
                self._allreduce(self.grads[o])
                L.check(eng.lib.raae_apply_adam(eng.handle, o, self.grads[o].data_ptr(), eng.stream))
        self._sync_bn_buffers()
        losses = torch.zeros(1, 12, dtype=torch.float32, device=eng.device)
        metrics = torch.zeros(1, 6, dtype=torch.float32, device=eng.device)
        L.check(eng.lib.raae_validate_epoch(eng.handle, int(epoch), losses.data_ptr(), metrics.data_ptr(), eng.stream))
        return losses[0], metrics[0]

    def train(self, max_epoch=None, callback=None):
        max_epoch = int(self.cfg["max_epoch"] if max_epoch is None else max_epoch)
        out = None
        for e in range(max_epoch):
            losses, metrics = self.train_epoch(e)
            out = metrics
            if callback is not None:
                callback(e, [float(v) for v in metrics[:5].cpu()])
        torch.cuda.synchronize(self.engine.device)
        return [float(v) for v in out[:5].cpu()]

    def state_vector(self):
        """Parameters of the three networks (for cross-rank consistency checks)."""
        lay, st = self.engine.lay, self.engine.state[0]
        return torch.cat([st[lay.net[i].param_off:lay.net[i].param_off + lay.net[i].n_params] for i in range(3)])
