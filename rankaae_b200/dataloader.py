"""Device-resident replacement of the reference's loader (`sc/clustering/dataloader.py:8-77`).

The CSV schema and the split rule are the reference's: two index columns, `AUX_*` x n_aux, then
`ENE_<energy>` x dim; rows are split sequentially (no shuffle of the split) by
`int(len * ratio)` with the remainder going to the test split (dataloader.py:12-25).  The file is
parsed ONCE (the reference parses it three times per trial) and each split becomes one float32
tensor; shuffling happens per epoch on the device (Engine.make_perm) instead of per-sample
`__getitem__` + collate on the host.
"""
import json
import os

import numpy as np
import pandas as pd


class AuxSpectraDataset:
    def __init__(self, csv_fn, split_portion, train_val_test_ratios=(0.7, 0.15, 0.15), n_aux=0, transform=None,
                 _full_df=None):
        self.metadata = {"path": csv_fn, "train_test_val_split_ratio": train_val_test_ratios}
        full_df = _full_df if _full_df is not None else pd.read_csv(csv_fn, index_col=[0, 1], comment='#')
        self.grid = np.array([float(col.strip('ENE_')) for col in full_df.columns if col.startswith('ENE_')])
        n_train_val_test = [int(len(full_df) * ratio) for ratio in train_val_test_ratios]
        n_train_val_test[-1] = int(len(full_df)) - sum(n_train_val_test[:-1])
        portion_options = ['train', 'val', 'test']
        assert split_portion in portion_options
        i_prev = portion_options.index(split_portion)
        df = full_df[sum(n_train_val_test[:i_prev]):sum(n_train_val_test[:i_prev + 1])]
        assert "ENE_" in df.columns.to_list()[n_aux]
        if n_aux > 0:
            assert "ENE_" not in df.columns.to_list()[n_aux - 1]
            assert "AUX_" in df.columns.to_list()[0]
            assert "AUX_" in df.columns.to_list()[n_aux - 1]
        data = df.to_numpy()
        self.spec = data[:, n_aux:]
        self.aux = data[:, :n_aux] if n_aux > 0 else None
        self.transform = transform
        self.atom_index = df.index.to_list()

    def __len__(self):
        return self.spec.shape[0]

    def __getitem__(self, idx):
        sample = (self.spec[idx], np.array([0.0])) if self.aux is None else (self.spec[idx], self.aux[idx])
        if self.transform is not None:
            sample = [self.transform(x) if x is not None else None for x in sample]
        return sample

    def tensors(self):
        """(spec float32 [n, dim], aux float32 [n, n_aux]) — the reference casts to float32 per sample
        (`torch.Tensor(sample)`, dataloader.py:61)."""
        spec = np.ascontiguousarray(self.spec, dtype=np.float32)
        aux = (np.ascontiguousarray(self.aux, dtype=np.float32) if self.aux is not None
               else np.zeros((len(self), 0), dtype=np.float32))
        return spec, aux


def get_datasets(csv_fn, train_val_test_ratios=(0.7, 0.15, 0.15), n_aux=0):
    """Parses the CSV once and returns the (train, val, test) splits."""
    full_df = pd.read_csv(csv_fn, index_col=[0, 1], comment='#')
    return [AuxSpectraDataset(csv_fn, p, train_val_test_ratios, n_aux=n_aux, _full_df=full_df)
            for p in ["train", "val", "test"]]


# ------------------------------------------------------------------------------------------
# loader front end for large sets and multi-GPU jobs: one parse per NODE, a binary cache next to the CSV
# (float32 .npy, memory-mapped by every rank), pinned staging to the device.  The reference parses the CSV three times per
# trial (dataloader.py:64-77), i.e. 24 times for 8 engines; at 1 M rows that alone costs minutes per trial.
# ------------------------------------------------------------------------------------------
CACHE_VERSION = 1


def _cache_paths(csv_fn, cache_dir=None):
    base = os.path.join(cache_dir, os.path.basename(csv_fn)) if cache_dir else csv_fn
    return base + ".raae.spec.npy", base + ".raae.aux.npy", base + ".raae.meta.json"


def _check_schema(columns, n_aux):
    """The reference's column asserts (dataloader.py:21-25), on the header."""
    assert "ENE_" in columns[n_aux]
    if n_aux > 0:
        assert "ENE_" not in columns[n_aux - 1]
        assert "AUX_" in columns[0]
        assert "AUX_" in columns[n_aux - 1]


def build_cache(csv_fn, n_aux, cache_dir=None, chunk_rows=100_000):
    """Parses the CSV once (in chunks, so a 1 M-row file never needs the float64 frame in memory) into float32 .npy files;
    written under temporary names and renamed, so a reader never sees a partial cache."""
    spec_p, aux_p, meta_p = _cache_paths(csv_fn, cache_dir)
    st = os.stat(csv_fn)
    specs, auxs, columns = [], [], None
    for df in pd.read_csv(csv_fn, index_col=[0, 1], comment='#', chunksize=chunk_rows):
        if columns is None:
            columns = df.columns.to_list()
            _check_schema(columns, n_aux)
        data = df.to_numpy()
        specs.append(np.ascontiguousarray(data[:, n_aux:], dtype=np.float32))
        auxs.append(np.ascontiguousarray(data[:, :n_aux], dtype=np.float32))
    spec = np.concatenate(specs) if specs else np.zeros((0, 0), np.float32)
    aux = np.concatenate(auxs) if auxs else np.zeros((0, n_aux), np.float32)
    for path, arr in ((spec_p, spec), (aux_p, aux)):
        tmp = path + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            np.save(f, arr)
        os.replace(tmp, path)
    meta = {"version": CACHE_VERSION, "csv_size": st.st_size, "csv_mtime_ns": st.st_mtime_ns, "n_aux": n_aux,
            "rows": int(spec.shape[0]), "dim": int(spec.shape[1]) if spec.ndim == 2 else 0,
            "grid": [float(c.strip('ENE_')) for c in columns if c.startswith('ENE_')]}
    tmp = meta_p + f".tmp{os.getpid()}"
    with open(tmp, "w") as f:
        json.dump(meta, f)
    os.replace(tmp, meta_p)
    return meta


def cache_valid(csv_fn, n_aux, cache_dir=None):
    spec_p, aux_p, meta_p = _cache_paths(csv_fn, cache_dir)
    try:
        meta = json.load(open(meta_p))
        st = os.stat(csv_fn)
        return (meta["version"] == CACHE_VERSION and meta["csv_size"] == st.st_size and meta["csv_mtime_ns"] == st.st_mtime_ns
                and meta["n_aux"] == n_aux and os.path.exists(spec_p) and os.path.exists(aux_p))
    except Exception:
        return False


def load_splits(csv_fn, train_val_test_ratios=(0.7, 0.15, 0.15), n_aux=0, rank=0, world=1, cache_dir=None, barrier=None):
    """[(spec, aux)] x (train, val, test) as float32 arrays memory-mapped from the binary cache.  Rank 0 parses the CSV when
    the cache is missing or stale; the other ranks wait at `barrier` (default: torch.distributed.barrier when a process
    group exists) and map the same files - one parse per node, whatever the number of GPUs and trials.  The split rule is
    the reference's (sequential, int(len * ratio), remainder to the test split; dataloader.py:14-20)."""
    if barrier is None:
        def barrier():
            import torch.distributed as dist
            if world > 1 and dist.is_available() and dist.is_initialized():
                dist.barrier()
    if rank == 0 and not cache_valid(csv_fn, n_aux, cache_dir):
        build_cache(csv_fn, n_aux, cache_dir)
    barrier()
    if not cache_valid(csv_fn, n_aux, cache_dir):
        raise RuntimeError(f"binary cache of {csv_fn} is missing or stale on rank {rank}")
    spec_p, aux_p, _ = _cache_paths(csv_fn, cache_dir)
    spec, aux = np.load(spec_p, mmap_mode="r"), np.load(aux_p, mmap_mode="r")
    n = spec.shape[0]
    counts = [int(n * r) for r in train_val_test_ratios]
    counts[-1] = n - sum(counts[:-1])
    out, lo = [], 0
    for c in counts:
        out.append((spec[lo:lo + c], aux[lo:lo + c]))
        lo += c
    return out


def to_device_pinned(arr, device, chunk_bytes=64 << 20):
    """Memory-mapped float32 array -> device tensor through a pinned staging buffer, chunk by chunk (two buffers in flight),
    so that a 1 GB split neither needs a second pageable copy nor a 1 GB pinned allocation."""
    import torch
    arr = np.asarray(arr) if not isinstance(arr, np.memmap) else arr
    n = arr.shape[0]
    out = torch.empty(arr.shape, dtype=torch.float32, device=device)
    if n == 0:
        return out
    row_bytes = max(1, int(np.prod(arr.shape[1:])) * 4)
    rows = max(1, chunk_bytes // row_bytes)
    pinned = torch.cuda.is_available() and torch.device(device).type == "cuda"
    bufs = [torch.empty((min(rows, n),) + tuple(arr.shape[1:]), dtype=torch.float32, pin_memory=pinned) for _ in range(2)]
    events = [None, None]
    for i, lo in enumerate(range(0, n, rows)):
        b = i & 1
        if events[b] is not None:
            events[b].synchronize()
        hi = min(n, lo + rows)
        bufs[b][:hi - lo].numpy()[...] = arr[lo:hi]           # straight from the page cache into the staging buffer
        out[lo:hi].copy_(bufs[b][:hi - lo], non_blocking=pinned)
        if pinned:
            events[b] = torch.cuda.Event()
            events[b].record()
    if pinned:
        torch.cuda.synchronize(device)
    return out
