"""Host-side owner of the device buffers behind the C ABI: per-trial state blocks, scratch,
hyper-parameters and the dataset.  PyTorch is used for device memory and streams only; every
number is produced by the kernels in csrc/ through librankaae_b200.so."""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib as L

ADAMW_DEFAULT_WD = 1e-2
# contractions on tcgen05 (3 x TF32 round-to-nearest split, fp32-grade: the parity suite passes in every mode): bit 0
# hidden-block forward, bit 1 hidden-block backward, bit 2 encoder input block from the batch / MI-phase operand images,
# bit 4 decoder output forward, bit 5 decoder output backward (the gradient tile parked in tensor memory); `tensor_cores: 0`
# in the config (or RAAE_TENSOR_CORES=0) selects the all-FP32-FMA path
DEFAULT_TENSOR_CORES = 55


def optimizer_hparams(cfg):
    """The five optimizers the gradient-reversal branch steps, Trainer.load_optimizers
    trainer.py:333-397 (phase order).  Returns (lr, beta1, beta2, weight_decay) per phase."""
    g = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    lr_base = g("lr_base", 1e-3)
    wd = g("weight_decay", 1e-2)
    dis_beta = g("dis_beta", 1.1)
    dflt = (0.9, 0.999)
    dis_betas = (dis_beta * 0.9, dis_beta * 0.009 + 0.99)
    if g("optimizer_name", "AdamW") == "Adam":
        raise NotImplementedError("only optimizer_name: AdamW is implemented (Adam's L2-in-gradient decay is not)")
    return [
        (g("lr_ratio_dis", 1) * lr_base, dis_betas[0], dis_betas[1], ADAMW_DEFAULT_WD),   # adversarial :380-387
        (g("lr_ratio_Corr", 1) * lr_base, dflt[0], dflt[1], wd),                          # correlation :357-363
        (g("lr_ratio_Reconn", 1) * lr_base, dflt[0], dflt[1], wd),                        # reconstruction :335-342
        (g("lr_ratio_Mutual", 1) * lr_base, dflt[0], dflt[1], ADAMW_DEFAULT_WD),          # mutual_info :343-349
        (g("lr_ratio_Smooth", 1) * lr_base, dflt[0], dflt[1], wd),                        # smoothness :350-356
    ]


def hp_row(cfg, seed=0):
    """One row of the float64 hyper-parameter table from fix_config.yaml keys."""
    g = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    hp = np.zeros(L.HP_COUNT, dtype=np.float64)
    for o, (lr, b1, b2, wd) in enumerate(optimizer_hparams(cfg)):
        hp[L.HP_LR0 + o], hp[L.HP_BETA1 + o], hp[L.HP_BETA2 + o], hp[L.HP_WD + o] = lr, b1, b2, wd
    hp[L.HP_DROPOUT] = g("dropout_rate", 0.2)
    hp[L.HP_DIS_DROPOUT] = g("dis_dropout_rate", 0.2)
    hp[L.HP_DIS_NOISE] = g("dis_noise", 0.1)
    hp[L.HP_SPEC_NOISE] = g("spec_noise", 0.0)
    hp[L.HP_ALPHA_FLAT_STEP] = g("alpha_flat_step", 800)
    hp[L.HP_ALPHA_LIMIT] = g("alpha_limit", 0.7)
    hp[L.HP_SCH_FACTOR] = g("sch_factor", 0.1)
    hp[L.HP_SCH_PATIENCE] = g("sch_patience", 100)
    hp[L.HP_EPOCH_STOP_SMOOTH] = g("epoch_stop_smooth", 500)      # trainer.py:59 default
    hp[L.HP_MAX_EPOCH] = g("max_epoch", 2000)
    hp[L.HP_SEED] = float(seed)
    return hp


def shapiro_weights(n):
    """Royston (AS R94) coefficients of scipy.stats.shapiro for sample size n, float64
    (SURVEY.md Appendix B); only the statistic is needed (trainer.py:287)."""
    from scipy.special import ndtri
    n2 = n // 2
    i = np.arange(1, n2 + 1, dtype=np.float64)
    m = ndtri((i - 0.375) / (n + 0.25))
    ssq = 2.0 * np.sum(m ** 2)
    r = 1.0 / math.sqrt(n)
    c1 = (0.0, 0.221157, -0.147981, -2.07119, 4.434685, -2.706056)
    c2 = (0.0, 0.042981, -0.293762, -1.752461, 5.682633, -3.582633)
    poly = lambda c: sum(ck * r ** k for k, ck in enumerate(c))
    a = np.zeros(n2)
    a1 = poly(c1) - m[0] / math.sqrt(ssq)
    if n > 5:
        a2 = poly(c2) - m[1] / math.sqrt(ssq)
        fac = math.sqrt((ssq - 2 * m[0] ** 2 - 2 * m[1] ** 2) / (1 - 2 * a1 ** 2 - 2 * a2 ** 2))
        a[0], a[1] = a1, a2
        a[2:] = -m[2:] / fac
    else:
        fac = math.sqrt((ssq - 2 * m[0] ** 2) / (1 - 2 * a1 ** 2))
        a[0] = a1
        a[1:] = -m[1:] / fac
    w = np.zeros(n)
    w[:n2] = -a
    w[n - n2:] = a[::-1]
    return w


def make_config(cfg, n_trials, max_rows=None):
    g = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    if g("ae_form", "FC") != "FC":
        raise NotImplementedError("only ae_form: FC is implemented by the fused path")
    if g("use_cnn_discriminator", False):
        raise NotImplementedError("only the FC discriminator is implemented by the fused path")
    if not g("gradient_reversal", True):
        raise NotImplementedError("only gradient_reversal: true is implemented by the fused path")
    act = g("decoder_activation", "ReLu")
    if act not in ("Softplus", "ReLu"):
        raise ValueError(f'Unknow activation function "{act}", please use one available in Pytorch')
    bs = int(g("batch_size", 1024))
    return L.Config(
        dim_in=int(g("dim_in", 256)), dim_out=int(g("dim_out", 256)), nstyle=int(g("nstyle", 5)),
        n_aux=int(g("n_aux", 0)), n_layers=int(g("n_layers", 3)), dis_layers=int(g("FC_discriminator_layers", 3)),
        batch_size=bs, n_trials=int(n_trials), kendall_activation=int(bool(g("kendall_activation", False))),
        use_flex_spec_target=int(bool(g("use_flex_spec_target", False))), decoder_softplus=int(act == "Softplus"),
        max_rows=int(max_rows if max_rows is not None else bs),
        ctas_per_trial=int(g("ctas_per_trial", int(os.environ.get("RAAE_CTAS_PER_TRIAL", "1")))),
        tensor_cores=int(g("tensor_cores", int(os.environ.get("RAAE_TENSOR_CORES", str(DEFAULT_TENSOR_CORES))))))


def auto_ctas_per_trial(cfg, n_trials, device=0):
    """Cluster size for `n_trials` resident trials when the config does not fix `ctas_per_trial`: the largest of 8 / 4 / 2
    whose clusters (one per trial) still fit the GPU in ONE wave (raae_max_clusters: 15 / 33 / 74 on a B200) and that has a
    128-row tile of the batch for every CTA; 1 (one CTA per trial, the ensemble path) otherwise."""
    g = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    fixed = g("ctas_per_trial", None)
    if fixed is None:
        fixed = os.environ.get("RAAE_CTAS_PER_TRIAL")
    if fixed is not None:
        return int(fixed)
    tiles = (int(g("batch_size", 1024)) + 127) // 128
    dev = torch.device(device).index if not isinstance(device, int) else device
    for c in (8, 4, 2):
        if tiles >= c and n_trials <= L.max_clusters(c, dev or 0):
            return c
    return 1


class Engine:
    """One handle per GPU: `n_trials` independent trials resident on `device`."""

    def __init__(self, cfg, n_trials=1, device="cuda:0", max_rows=None, seeds=None, per_trial_cfg=None):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.RaaeError("rankaae_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device)
        self.cfg = cfg
        self.ccfg = make_config(cfg, n_trials, max_rows)
        self.n_trials = int(n_trials)
        self.lay = L.query_layout(self.ccfg)
        self.handle = L._p()
        L.check(self.lib.raae_create(C.byref(self.ccfg), self.device.index or 0, C.byref(self.handle)))
        self.state = torch.zeros(self.n_trials, self.lay.state_floats, dtype=torch.float32, device=self.device)
        self.scratch = torch.zeros(self.n_trials, self.lay.scratch_floats, dtype=torch.float32, device=self.device)
        seeds = list(range(self.n_trials)) if seeds is None else list(seeds)
        rows = [hp_row(per_trial_cfg[t] if per_trial_cfg is not None else cfg, seeds[t]) for t in range(self.n_trials)]
        self.hp = torch.from_numpy(np.stack(rows)).to(self.device)
        L.check(self.lib.raae_bind_state(self.handle, self.state.data_ptr(), self.scratch.data_ptr(), self.hp.data_ptr()))
        self._keep = []          # tensors whose pointers the library borrows
        self.n_train = self.n_val = 0
        self.reset_optimizers()

    def close(self):
        if self.handle:
            self.lib.raae_destroy(self.handle)
            self.handle = L._p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset_optimizers(self):
        L.check(self.lib.raae_reset_optimizers(self.handle, self.stream))

    def set_hp(self, trial, hp_np):
        self.hp[trial].copy_(torch.from_numpy(np.asarray(hp_np, dtype=np.float64)))

    def bind_dataset(self, spec_train, aux_train, spec_val, aux_val, rows_per_trial=None):
        """float32 row-major device tensors (copied to the device if needed).  `rows_per_trial` (data-parallel replicas,
        dp.py): every trial works on its own `rows_per_trial` rows - the permutations handed to the kernels have that
        length and index the whole bound array."""
        def dev(x):
            return torch.as_tensor(x, dtype=torch.float32).to(self.device).contiguous()
        st, at, sv, av = dev(spec_train), dev(aux_train), dev(spec_val), dev(aux_val)
        self._data = (st, at, sv, av)
        self.n_train, self.n_val = st.shape[0] if rows_per_trial is None else int(rows_per_trial), sv.shape[0]
        assert self.n_train <= st.shape[0]
        L.check(self.lib.raae_bind_dataset(self.handle, st.data_ptr(), at.data_ptr(), self.n_train,
                                           sv.data_ptr(), av.data_ptr(), self.n_val))
        if self.n_val >= 3:
            self._shapiro = torch.from_numpy(shapiro_weights(self.n_val).astype(np.float32)).to(self.device)
            L.check(self.lib.raae_bind_shapiro_weights(self.handle, self._shapiro.data_ptr(), self.n_val))

    # ------------------------------------------------------------------ state <-> numpy / modules
    def _net_slices(self, net_i):
        """[(key, layer, offset_in_state, shape)] in parameters() order."""
        n = self.lay.net[net_i]
        out = []
        for l in range(n.n_linear):
            out.append(("W", l, n.param_off + n.w_off[l], (n.out_dim[l], n.in_dim[l])))
            out.append(("b", l, n.param_off + n.b_off[l], (n.out_dim[l],)))
            if n.a_off[l] >= 0:
                out.append(("a", l, n.param_off + n.a_off[l], (n.out_dim[l],)))
        return out

    def set_state(self, trial, state, opt=None):
        """Loads an oracle-format state ({net: {W, b, a, rm, rv, nbt}}) and optionally the AdamW state
        ({phase: {t, lr, m: {net: ...}, v: ...}}) into trial `trial`."""
        blk = np.zeros(self.lay.state_floats, dtype=np.float32)
        blk[:] = self.state[trial].cpu().numpy()
        for ni, net in enumerate(L.NETS):
            n = self.lay.net[ni]
            for key, l, off, shp in self._net_slices(ni):
                blk[off:off + int(np.prod(shp))] = np.asarray(state[net][key][l], dtype=np.float32).ravel()
            if net != "S":
                for l in range(len(state[net]["rm"])):
                    blk[n.rm_off[l]:n.rm_off[l] + n.out_dim[l]] = state[net]["rm"][l]
                    blk[n.rv_off[l]:n.rv_off[l] + n.out_dim[l]] = state[net]["rv"][l]
                blk[n.nbt_off] = float(state[net].get("nbt", 0))
        if opt is not None:
            for o, ph in enumerate(L.PHASES):
                ol = self.lay.opt[o]
                blk[ol.scalar_off + 0] = opt[ph]["lr"]
                blk[ol.scalar_off + 1] = float(opt[ph]["t"])
                for ni, net in enumerate(L.NETS):
                    if ol.net_off[ni] < 0:
                        continue
                    base = self.lay.net[ni].param_off
                    for key, l, off, shp in self._net_slices(ni):
                        rel = off - base
                        cnt = int(np.prod(shp))
                        blk[ol.m_off + ol.net_off[ni] + rel:ol.m_off + ol.net_off[ni] + rel + cnt] = \
                            np.asarray(opt[ph]["m"][net][key][l], dtype=np.float32).ravel()
                        blk[ol.v_off + ol.net_off[ni] + rel:ol.v_off + ol.net_off[ni] + rel + cnt] = \
                            np.asarray(opt[ph]["v"][net][key][l], dtype=np.float32).ravel()
        self.state[trial].copy_(torch.from_numpy(blk))

    def get_state(self, trial):
        """Returns (state, opt) in the oracle format (float32 numpy)."""
        blk = self.state[trial].cpu().numpy()
        state, opt = {}, {}
        for ni, net in enumerate(L.NETS):
            n = self.lay.net[ni]
            d = {"W": [], "b": [], "a": []}
            for key, l, off, shp in self._net_slices(ni):
                d[key].append(blk[off:off + int(np.prod(shp))].reshape(shp).copy())
            if net != "S":
                nbn = n.n_linear if net == "E" else n.n_linear - 1
                d["rm"] = [blk[n.rm_off[l]:n.rm_off[l] + n.out_dim[l]].copy() for l in range(nbn)]
                d["rv"] = [blk[n.rv_off[l]:n.rv_off[l] + n.out_dim[l]].copy() for l in range(nbn)]
                d["nbt"] = int(round(float(blk[n.nbt_off])))
            state[net] = d
        for o, ph in enumerate(L.PHASES):
            ol = self.lay.opt[o]
            od = {"lr": float(blk[ol.scalar_off]), "t": int(round(float(blk[ol.scalar_off + 1]))),
                  "best": float(blk[ol.scalar_off + 2]), "bad": float(blk[ol.scalar_off + 3]), "m": {}, "v": {}}
            for ni, net in enumerate(L.NETS):
                if ol.net_off[ni] < 0:
                    continue
                base = self.lay.net[ni].param_off
                md, vd = {"W": [], "b": [], "a": []}, {"W": [], "b": [], "a": []}
                for key, l, off, shp in self._net_slices(ni):
                    rel, cnt = off - base, int(np.prod(shp))
                    md[key].append(blk[ol.m_off + ol.net_off[ni] + rel:][:cnt].reshape(shp).copy())
                    vd[key].append(blk[ol.v_off + ol.net_off[ni] + rel:][:cnt].reshape(shp).copy())
                od["m"][net], od["v"][net] = md, vd
            opt[ph] = od
        return state, opt

    def split_grad_vector(self, phase_i, vec):
        """Optimizer-order gradient vector -> {net: {W, b, a}}."""
        ol = self.lay.opt[phase_i]
        out = {}
        for ni, net in enumerate(L.NETS):
            if ol.net_off[ni] < 0:
                continue
            base = self.lay.net[ni].param_off
            d = {"W": [], "b": [], "a": []}
            for key, l, off, shp in self._net_slices(ni):
                rel, cnt = off - base, int(np.prod(shp))
                d[key].append(vec[ol.net_off[ni] + rel:ol.net_off[ni] + rel + cnt].reshape(shp).copy())
            out[net] = d
        return out

    @staticmethod
    def _linear_stack(seq):
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        pre = [m for m in seq if isinstance(m, torch.nn.PReLU)]
        bns = [m for m in seq if isinstance(m, torch.nn.BatchNorm1d)]
        return lin, pre, bns

    def load_modules(self, trial, encoder, decoder, discriminator):
        """nn.Module parameters / BN buffers -> state block (AdamW moments untouched)."""
        state = {}
        for net, mod in (("E", encoder), ("D", decoder), ("S", discriminator)):
            lin, pre, bns = self._linear_stack(mod.main)
            d = {"W": [m.weight.detach().cpu().numpy() for m in lin], "b": [m.bias.detach().cpu().numpy() for m in lin],
                 "a": [m.weight.detach().cpu().numpy() for m in pre]}
            if net != "S":
                d["rm"] = [m.running_mean.detach().cpu().numpy() for m in bns]
                d["rv"] = [m.running_var.detach().cpu().numpy() for m in bns]
                d["nbt"] = int(bns[0].num_batches_tracked) if bns else 0
            state[net] = d
        self.set_state(trial, state)

    def store_modules(self, trial, encoder, decoder, discriminator):
        """state block -> nn.Module parameters / BN buffers (what final.pt pickles, trainer.py:281-283)."""
        state, _ = self.get_state(trial)
        with torch.no_grad():
            for net, mod in (("E", encoder), ("D", decoder), ("S", discriminator)):
                lin, pre, bns = self._linear_stack(mod.main)
                for l, m in enumerate(lin):
                    m.weight.copy_(torch.from_numpy(state[net]["W"][l]))
                    m.bias.copy_(torch.from_numpy(state[net]["b"][l]))
                for l, m in enumerate(pre):
                    m.weight.copy_(torch.from_numpy(state[net]["a"][l]))
                if net != "S":
                    for l, m in enumerate(bns):
                        m.running_mean.copy_(torch.from_numpy(state[net]["rm"][l]))
                        m.running_var.copy_(torch.from_numpy(state[net]["rv"][l]))
                        m.num_batches_tracked.fill_(state[net]["nbt"])

    # ------------------------------------------------------------------ compute entry points
    def _f32(self, x):
        if x is None:
            return None
        t = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).to(self.device).contiguous()
        self._keep.append(t)
        return t

    def _u8(self, x):
        if x is None:
            return None
        t = torch.as_tensor(np.ascontiguousarray(x).astype(np.uint8)).to(self.device).contiguous()
        self._keep.append(t)
        return t

    def step_debug(self, trial, x_noisy, aux, rnd=None, epoch=0, phase_mask=0x1f, apply_updates=True,
                   want_grads=True):
        """One teacher-forced step (trainer.py:112-204) with explicit random draws `rnd` in the
        oracle's naming (E0..E5, D0..D3, S_real_masks, S_fake_masks, z_real, S_real_eps, S_fake_eps,
        z_sample); missing entries are drawn by the in-kernel generator."""
        rnd = rnd or {}
        self._keep = []
        io = L.DebugIO()
        xs, au = self._f32(x_noisy), self._f32(aux)
        io.x_noisy, io.aux = xs.data_ptr(), au.data_ptr()
        io.rows, io.epoch, io.phase_mask, io.apply_updates = xs.shape[0], int(epoch), int(phase_mask), int(bool(apply_updates))
        for i in range(6):
            for l, m in enumerate(rnd.get(f"E{i}") or []):
                io.mask_enc[i][l] = self._u8(m).data_ptr()
        for i in range(4):
            for l, m in enumerate(rnd.get(f"D{i}") or []):
                io.mask_dec[i][l] = self._u8(m).data_ptr()
        for i, key in enumerate(("S_real_masks", "S_fake_masks")):
            for l, m in enumerate(rnd.get(key) or []):
                io.mask_dis[i][l] = self._u8(m).data_ptr()
        for field, key in (("z_real", "z_real"), ("dis_eps_real", "S_real_eps"), ("dis_eps_fake", "S_fake_eps"),
                           ("z_sample", "z_sample")):
            t = self._f32(rnd.get(key))
            if t is not None:
                setattr(io, field, t.data_ptr())
        losses = torch.zeros(L.NUM_PHASES, dtype=torch.float32, device=self.device)
        styles = torch.zeros(xs.shape[0], self.ccfg.nstyle, dtype=torch.float32, device=self.device)
        io.losses, io.styles = losses.data_ptr(), styles.data_ptr()
        gvecs = []
        if want_grads:
            for o in range(L.NUM_PHASES):
                gv = torch.zeros(self.lay.opt[o].n, dtype=torch.float32, device=self.device)
                gvecs.append(gv)
                io.grads[o] = gv.data_ptr()
        L.check(self.lib.raae_step_debug(self.handle, trial, C.byref(io), self.stream))
        torch.cuda.synchronize(self.device)
        out = {"losses": dict(zip(L.PHASES, losses.cpu().numpy().astype(np.float64))), "styles": styles.cpu().numpy()}
        if want_grads:
            out["grads"] = {ph: self.split_grad_vector(o, gvecs[o].cpu().numpy()) for o, ph in enumerate(L.PHASES)}
        self._keep = []
        return out

    def validate(self, trial, z_sample=None, z_real=None, epoch=0, avg_mutual_info=None):
        """The eval block (trainer.py:207-297) of one trial with optional explicit draws."""
        self._keep = []
        io = L.ValIO()
        zs, zr = self._f32(z_sample), self._f32(z_real)
        if zs is not None:
            io.z_sample = zs.data_ptr()
        if zr is not None:
            io.z_real = zr.data_ptr()
        io.epoch = int(epoch)
        io.avg_mutual_info = float("-inf") if avg_mutual_info is None else float(avg_mutual_info)
        losses = torch.zeros(L.NUM_PHASES, dtype=torch.float32, device=self.device)
        metrics = torch.zeros(6, dtype=torch.float32, device=self.device)
        z = torch.zeros(self.n_val, self.ccfg.nstyle, dtype=torch.float32, device=self.device)
        io.losses, io.metrics, io.z = losses.data_ptr(), metrics.data_ptr(), z.data_ptr()
        L.check(self.lib.raae_validate(self.handle, trial, C.byref(io), self.stream))
        torch.cuda.synchronize(self.device)
        self._keep = []
        return {"losses": dict(zip(L.PHASES, losses.cpu().numpy().astype(np.float64))),
                "metrics": metrics.cpu().numpy().astype(np.float64), "z": z.cpu().numpy()}

    def make_perm(self, n_epochs, generator=None):
        """[n_epochs][n_trials][n_train] int32 shuffles (the DataLoader's RandomSampler order,
        dataloader.py:70-71), drawn on the device."""
        r = torch.rand(n_epochs, self.n_trials, self.n_train, device=self.device, generator=generator)
        return torch.argsort(r, dim=-1).to(torch.int32).contiguous()

    def train_epochs(self, epoch_begin, n_epochs, perm=None):
        """Runs epochs [epoch_begin, epoch_begin + n_epochs) for every resident trial; returns device tensors
        losses [n_epochs][n_trials][12] and metrics [n_epochs][n_trials][6] (not synchronised)."""
        if perm is None:
            perm = self.make_perm(n_epochs)
        assert perm.dtype == torch.int32 and tuple(perm.shape) == (n_epochs, self.n_trials, self.n_train)
        losses = torch.zeros(n_epochs, self.n_trials, 12, dtype=torch.float32, device=self.device)
        metrics = torch.zeros(n_epochs, self.n_trials, 6, dtype=torch.float32, device=self.device)
        L.check(self.lib.raae_train_epochs(self.handle, int(epoch_begin), int(n_epochs), perm.data_ptr(),
                                           losses.data_ptr(), metrics.data_ptr(), self.stream))
        self._perm = perm
        return losses, metrics

    @property
    def launch_count(self):
        return int(self.lib.raae_launch_count(self.handle))
