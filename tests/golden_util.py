"""Loads tests/golden/*.npz (written by oracle/make_golden.py from the unmodified reference)
into the structures the oracle and the CUDA host wrapper use."""
import ast
import os

import numpy as np

from oracle import aae_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NETS = ("E", "D", "S")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name))
        self.cfg_dict = {k: ast.literal_eval(v) for k, v in zip(self.z["config_keys"], self.z["config_vals"])}
        self.cfg = O.Config.from_dict(self.cfg_dict)
        self.record_epoch = int(self.z["record_epoch"])
        self.spec = self.z["spec"].astype(np.float64)
        self.aux = self.z["aux"].astype(np.float64)
        n = self.spec.shape[0]
        self.n_train = int(n * 0.7)
        self.n_val = int(n * 0.15)
        self.n_batches = (self.n_train + self.cfg.batch_size - 1) // self.cfg.batch_size

    def state(self, prefix):
        st = {}
        for net in NETS:
            d = {}
            for k in ("W", "b", "a", "rm", "rv"):
                lst = []
                i = 0
                while f"{prefix}.{net}.{k}{i}" in self.z:
                    lst.append(self.z[f"{prefix}.{net}.{k}{i}"].astype(np.float64))
                    i += 1
                if lst or k in ("W", "b", "a"):
                    d[k] = lst
            if f"{prefix}.{net}.nbt" in self.z:
                d["nbt"] = int(self.z[f"{prefix}.{net}.nbt"])
            st[net] = d
        st["S"].pop("rm", None)
        st["S"].pop("rv", None)
        return st

    def opt(self, prefix="opt0"):
        opt = {}
        hp = self.cfg.optimizer_hparams()
        for name in O.PHASES:
            o = dict(t=int(self.z[f"{prefix}.{name}.t"]), lr=float(self.z[f"lr0.{name}"]), m={}, v={})
            for mv in ("m", "v"):
                for net in hp[name]["nets"]:
                    d = {}
                    for k in ("W", "b", "a"):
                        lst = []
                        i = 0
                        while f"{prefix}.{name}.{mv}.{net}.{k}{i}" in self.z:
                            lst.append(self.z[f"{prefix}.{name}.{mv}.{net}.{k}{i}"].astype(np.float64))
                            i += 1
                        d[k] = lst
                    o[mv][net] = d
            opt[name] = o
        return opt

    def batch(self, b):
        """Returns (x_noisy, aux, rnd) for batch b exactly as the reference consumed them."""
        pre = f"b{b}"
        idx = self.z[f"{pre}.idx"]
        eps = self.z[f"{pre}.eps_x"].astype(np.float64)
        x_noisy = self.spec[idx] + eps * self.cfg.spec_noise
        aux = self.aux[idx]

        def masks(key, n):
            out = []
            for i in range(n):
                shp = tuple(self.z[f"{pre}.maskshape.{key}.{i}"])
                bits = np.unpackbits(self.z[f"{pre}.mask.{key}.{i}"])[: shp[0] * shp[1]]
                out.append(bits.reshape(shp).astype(np.float64))
            return out

        L1 = self.cfg.n_layers - 1
        nS = self.cfg.dis_layers - 1
        rnd = {}
        for i in range(6):
            rnd[f"E{i}"] = masks(f"E{i}", L1)
        for i in range(4):
            rnd[f"D{i}"] = masks(f"D{i}", L1)
        rnd["S_real_masks"] = masks("S_real", nS)
        rnd["S_fake_masks"] = masks("S_fake", nS)
        for k in ("z_real", "S_real_eps", "S_fake_eps", "z_sample"):
            rnd[k] = self.z[f"{pre}.{k}"].astype(np.float64)
        return x_noisy, aux, rnd

    def losses(self, b):
        return dict(zip(O.PHASES, self.z[f"b{b}.losses"]))

    def grad(self, b, phase, net, key):
        name = f"b{b}.grad.{phase}.{net}.{key}"
        return self.z[name].astype(np.float64) if name in self.z else None

    def gradsum(self, b, phase, net, key):
        name = f"b{b}.gradsum.{phase}.{net}.{key}"
        return self.z[name] if name in self.z else None

    def val(self):
        n0 = self.n_train
        return dict(spec=self.spec[n0:n0 + self.n_val], aux=self.aux[n0:n0 + self.n_val],
                    z_sample=self.z["val.z_sample"].astype(np.float64),
                    z_real=self.z["val.z_real"].astype(np.float64),
                    losses=dict(zip(O.PHASES, self.z["val.losses"])),
                    metrics=self.z["val.metrics"])


def iter_params(state_or_grads, nets=NETS):
    for net in nets:
        if net not in state_or_grads:
            continue
        for k in ("W", "b", "a"):
            for i, x in enumerate(state_or_grads[net][k]):
                yield net, f"{k}{i}", x


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
