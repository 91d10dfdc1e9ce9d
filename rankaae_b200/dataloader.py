"""Device-resident replacement of the reference's loader (`sc/clustering/dataloader.py:8-77`).

The CSV schema and the split rule are the reference's: two index columns, `AUX_*` x n_aux, then
`ENE_<energy>` x dim; rows are split sequentially (no shuffle of the split) by
`int(len * ratio)` with the remainder going to the test split (dataloader.py:12-25).  The file is
parsed ONCE (the reference parses it three times per trial) and each split becomes one float32
tensor; shuffling happens per epoch on the device (Engine.make_perm) instead of per-sample
`__getitem__` + collate on the host.
"""
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
