"""Single-trial data-parallel training (BASELINE.json configs[3]: one trial, per-rank batch, gradient all-reduce).

Every rank holds the same weights and AdamW state and its own shard of the training rows.  Each loss phase of each batch
is one split-phase kernel launch that exports the phase's gradient vector instead of applying it
(`raae_train_phase`), followed by the exchange of that vector (118-238 KB, 5 per step, latency-bound):

* `exchange="peer"` (default for world > 1): ONE launch, `raae_apply_adam_peer` — the ranks signal each other through
  flag words in peer-mapped memory, every rank sums the gradient vectors in rank order straight out of the peers' HBM
  (P2P loads over NVLink, CUDA IPC mappings, no NCCL call), divides by world and applies AdamW in the same kernel;
* `exchange="nccl"`: `torch.distributed.all_reduce` + divide + `raae_apply_adam` (+ its step-counter kernel) — four
  launches, kept as the yardstick (bit-identical to the peer path for world 2, where the sum has one order).

BatchNorm batch statistics and Kendall pairs stay rank-local (DistributedDataParallel semantics, not large-batch
semantics); the BatchNorm running buffers are averaged across ranks before every validation block so that all ranks
score the same model.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from .engine import Engine, auto_ctas_per_trial
from .trainer import build_modules


def shard_rows(n_rows, world, rank):
    """Contiguous, equally sized shards (the remainder rows are dropped so that every rank runs the same number of
    batches — a collective per phase needs that)."""
    per = n_rows // world
    return rank * per, (rank + 1) * per


def shard_layout(n_rows, world, rank, replicas=1):
    """Rows of a rank under `world x replicas` equally sized shards (shard q = rank * replicas + replica; remainder rows
    dropped): returns (rows per shard, first row of the rank, one past its last row)."""
    per = n_rows // (world * replicas)
    return per, rank * replicas * per, (rank + 1) * replicas * per


class DataParallelTrainer:
    """`replicas` > 1 puts that many data-parallel replicas of the trial on EVERY GPU (one CTA / SM each, identical
    weights, own shard, own noise streams): the job then has world x replicas shards with DistributedDataParallel
    semantics - global batch = batch_size x world x replicas - and the exchange kernel averages over all of them.
    replicas = 1 is BASELINE.json configs[3] as written (one 512-row batch per GPU, one SM busy per GPU)."""

    def __init__(self, cfg, spec_train, aux_train, spec_val, aux_val, device, rank=0, world=1, seed=0, exchange=None,
                 replicas=1, shard_seeds=None, presharded=False):
        """`presharded`: spec_train / aux_train are already THIS rank's rows (e.g. generated or loaded per rank); otherwise
        they are the whole training split and the rank takes its contiguous share (shard_layout)."""
        self.rank, self.world, self.replicas = rank, world, int(replicas)
        V = self.replicas
        if exchange is None:
            exchange = os.environ.get("RAAE_DP_EXCHANGE", "peer" if world * V > 1 else "nccl")
        if exchange not in ("peer", "nccl"):
            raise ValueError(f"exchange must be 'peer' or 'nccl', got {exchange!r}")
        if world > L.MAX_PEERS and exchange == "peer":
            raise ValueError(f"the peer exchange supports up to {L.MAX_PEERS} ranks (one NVLink domain)")
        self.exchange = exchange
        self.cfg = dict(cfg)
        self.cfg.setdefault("epoch_stop_smooth", 500)
        self.cfg["ctas_per_trial"] = auto_ctas_per_trial(self.cfg, V, device)     # per-GPU batch 512: a 4-CTA cluster per replica
        if presharded:
            self.per, lo, hi = len(spec_train) // V, 0, (len(spec_train) // V) * V
        else:
            self.per, lo, hi = shard_layout(len(spec_train), world, rank, V)  # rows of one shard; shard q = rank * V + replica
        self.engine = Engine(self.cfg, n_trials=V, device=device, max_rows=max(int(cfg["batch_size"]), len(spec_val)),
                             seeds=shard_seeds if shard_seeds is not None else
                             [seed * 1000 + rank * V + v for v in range(V)])         # shard-local noise / dropout streams
        self.modules = build_modules(self.cfg, seed=seed)             # identical initial weights on every rank / replica
        for v in range(V):
            self.engine.load_modules(v, *self.modules)
        self.engine.bind_dataset(spec_train[lo:hi], aux_train[lo:hi], spec_val, aux_val, rows_per_trial=self.per)
        lay = self.engine.lay
        self._gptr = (L._p * L.NUM_PHASES)()
        if exchange == "peer":
            self._connect_peers()
        else:
            self.grads = [torch.zeros(V, lay.opt[o].n, dtype=torch.float32, device=self.engine.device) for o in range(L.NUM_PHASES)]
            self._grad_ptrs = [g.data_ptr() for g in self.grads]
        bs = int(cfg["batch_size"])
        self.n_steps = (self.per + bs - 1) // bs
        self._shard_off = (torch.arange(V, device=self.engine.device, dtype=torch.int32) * self.per).view(V, 1)

    def make_perm(self):
        """One shuffle per replica, as indices into the rank's rows: replica v owns rows [v * per, (v + 1) * per)."""
        return (self.engine.make_perm(1)[0] + self._shard_off).contiguous()

    def _connect_peers(self):
        """Allocates this rank's exchange block, gathers the CUDA IPC handles of all ranks and maps their blocks."""
        eng = self.engine
        mine = C.create_string_buffer(L.IPC_HANDLE_BYTES)
        L.check(eng.lib.raae_peer_alloc(eng.handle, self.world, self.rank, mine))
        handles = bytes(mine.raw)
        if self.world > 1:                              # [world][64] bytes in rank order, gathered on this rank's own device
            t = torch.frombuffer(bytearray(handles), dtype=torch.uint8).to(eng.device)
            out = torch.empty(self.world * L.IPC_HANDLE_BYTES, dtype=torch.uint8, device=eng.device)
            torch.distributed.all_gather_into_tensor(out, t)
            handles = out.cpu().numpy().tobytes()
        L.check(eng.lib.raae_peer_connect(eng.handle, handles))
        self._grad_ptrs = []
        for o in range(L.NUM_PHASES):
            ptr = L._p()
            L.check(eng.lib.raae_peer_grad_ptr(eng.handle, o, C.byref(ptr)))
            self._grad_ptrs.append(ptr.value)
        if self.world > 1:                              # every block is mapped before the first flag is written
            torch.distributed.all_reduce(torch.zeros(1, device=eng.device))
            torch.cuda.synchronize(eng.device)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        # a trainer dropped without close(): still barrier + unmap before the engine frees the exported block, so that no
        # peer polls or reads freed memory (skipped when the process group is already gone)
        try:
            if self.world == 1 or torch.distributed.is_initialized():
                self.close()
        except Exception:
            pass

    def close(self):
        """Barrier + unmap (no rank may free its block while a peer can still read it), then the engine."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        if self.exchange == "peer" and self.engine.handle:
            torch.cuda.synchronize(self.engine.device)
            if self.world > 1:
                torch.distributed.all_reduce(torch.zeros(1, device=self.engine.device))
                torch.cuda.synchronize(self.engine.device)
            L.check(self.engine.lib.raae_peer_free(self.engine.handle))
        self.engine.close()

    def _allreduce(self, t):
        if self.world > 1:
            torch.distributed.all_reduce(t)
            t /= self.world

    def _mean_over_shards(self, block):
        """block: [replicas][k] view of the state -> every replica (on every rank) gets the mean over all shards."""
        m = block.mean(0, keepdim=True) if self.replicas > 1 else block.clone()
        self._allreduce(m)
        block.copy_(m.expand_as(block))

    def _sync_bn_buffers(self):
        """Before the validation block: BatchNorm running buffers and the running sum / count of the mutual-information
        loss (the one rank-local input of the combined metric, trainer.py:294-297) are averaged over all shards, so every
        replica scores the same model with the same metric and the ReduceLROnPlateau states cannot drift apart."""
        if self.world * self.replicas == 1:
            return
        lay, st = self.engine.lay, self.engine.state
        for ni in (0, 1):
            n = lay.net[ni]
            nbn = n.n_linear if ni == 0 else n.n_linear - 1
            lo, hi = n.rm_off[0], n.rv_off[nbn - 1] + n.out_dim[nbn - 1]
            self._mean_over_shards(st[:, lo:hi])
        self._mean_over_shards(st[:, lay.misc_off + 5:lay.misc_off + 7])

    def _exchange_update(self, o):
        eng = self.engine
        if self.exchange == "peer":
            L.check(eng.lib.raae_apply_adam_peer(eng.handle, o, eng.stream))
        else:
            g = self.grads[o]
            if self.replicas > 1:
                m = g.sum(0, keepdim=True)
                self._allreduce(m)
                g.copy_((m / self.replicas).expand_as(g))
            else:
                self._allreduce(g)
            L.check(eng.lib.raae_apply_adam(eng.handle, o, self._grad_ptrs[o], eng.stream))

    def train_epoch(self, epoch, perm=None):
        """One epoch: every batch runs its five phases as launch -> gradient exchange + AdamW.  `perm`: int32
        [replicas][rows per shard] indices into the rank's rows (make_perm).  Returns (losses[12], metrics[6])."""
        eng = self.engine
        if perm is None:
            perm = self.make_perm()
        perm = perm.view(self.replicas, self.per)
        stop_smooth = float(self.cfg["epoch_stop_smooth"])
        for s in range(self.n_steps):
            for o in range(L.NUM_PHASES):
                if o == 4 and epoch >= stop_smooth:
                    continue
                for k in range(L.NUM_PHASES):
                    self._gptr[k] = self._grad_ptrs[k] if k == o else None
                L.check(eng.lib.raae_train_phase(eng.handle, int(epoch), s, 1 << o, perm.data_ptr(), self._gptr, eng.stream))
                self._exchange_update(o)
        self._sync_bn_buffers()
        losses = torch.zeros(self.replicas, 12, dtype=torch.float32, device=eng.device)
        metrics = torch.zeros(self.replicas, 6, dtype=torch.float32, device=eng.device)
        L.check(eng.lib.raae_validate_epoch(eng.handle, int(epoch), losses.data_ptr(), metrics.data_ptr(), eng.stream))
        return losses[0], metrics[0]

    def train(self, max_epoch=None, callback=None):
        if max_epoch is not None and int(max_epoch) != int(self.cfg["max_epoch"]):
            # the in-kernel alpha schedule divides by the hp row's max_epoch (functions.py:214-219): keep it in step
            self.engine.hp[:, L.HP_MAX_EPOCH] = float(max_epoch)
            self.cfg["max_epoch"] = int(max_epoch)
        max_epoch = int(self.cfg["max_epoch"])
        out = None
        for e in range(max_epoch):
            losses, metrics = self.train_epoch(e)
            out = metrics
            if callback is not None:
                callback(e, [float(v) for v in metrics[:5].cpu()])
                if self.world > 1 and self.exchange == "peer":
                    # a slow callback on one rank (checkpoint I/O) must not run into the exchange kernel's arrival timeout
                    torch.distributed.barrier()
        torch.cuda.synchronize(self.engine.device)
        self.check_exchange()
        return [float(v) for v in out[:5].cpu()]

    def check_exchange(self):
        """Raises if an exchange gave up waiting for a peer (the kernel skips that update instead of trapping; the job is
        inconsistent from then on).  Synchronises the device."""
        if self.exchange != "peer":
            return
        failed = C.c_uint(0)
        L.check(self.engine.lib.raae_peer_status(self.engine.handle, C.byref(failed)))
        if failed.value:
            raise L.RaaeError(f"peer exchange #{failed.value} timed out waiting for a rank (RAAE_PEER_TIMEOUT_S)")

    def state_vector(self):
        """Parameters of the three networks (for cross-rank consistency checks)."""
        lay, st = self.engine.lay, self.engine.state[0]
        return torch.cat([st[lay.net[i].param_off:lay.net[i].param_off + lay.net[i].n_params] for i in range(3)])
