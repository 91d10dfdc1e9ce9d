"""Generates tests/golden/*.npz by running the UNMODIFIED reference trainer
(`/root/reference/sc/clustering/trainer.py`, `Trainer.from_data(...).train()`) in float64 on a
small synthetic CSV, with every random draw recorded (TEST INFRASTRUCTURE ONLY).

    python oracle/make_golden.py            # writes tests/golden/step_fresh.npz, step_warm.npz

What is patched, and why it does not change the arithmetic:
  * torch.randn / torch.randn_like / torch.nn.functional.dropout are wrapped so that each draw
    is made by torch itself, rounded to a float32-representable value, logged, and applied with
    the stock formula (x * mask / (1 - p));
  * every optimizer's .step is wrapped to snapshot the .grad tensors just before it runs;
  * the five loss functions in the trainer's namespace are wrapped to log their return values;
  * the explicit-float32 BCE labels of adversarial_loss are promoted to float64 (values 0/1);
  * GaussianSmoothing's float32 taps are cast to the activation dtype (values unchanged);
  * torch default dtype is float64 so that `torch.Tensor(sample)` (dataloader.py:61), the
    modules and the draws are all double.
At the start of the recorded epoch all parameters, BN buffers and Adam moments are rounded to
float32-representable values (in place, still float64), so the float32 arrays stored in the
fixture ARE the exact inputs of the float64 run.

Runs only where /root/reference exists.  The fixtures are committed; tests never need this file.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import aae_oracle as O  # noqa: E402

CONFIG = dict(
    data_file="synthetic.csv", trials=1, timeout=10, verbose=False, max_epoch=1, batch_size=200,
    gradient_reversal=True, alpha_flat_step=739, alpha_limit=0.7172, decoder_activation="Softplus",
    dis_beta=1.1, dis_dropout_rate=0.056, dis_noise=0.56, gen_beta=1.1,
    n_aux=5, nstyle=6, ae_form="FC", dim_in=256, dim_out=256, n_layers=5, FC_discriminator_layers=3,
    use_cnn_discriminator=False, dropout_rate=0.04, sch_factor=0.1, sch_patience=100,
    lr_base=0.001, lr_ratio_Corr=10, lr_ratio_Mutual=1, lr_ratio_Reconn=10, lr_ratio_Smooth=1,
    lr_ratio_dis=1, lr_ratio_gen=10, optimizer_name="AdamW", spec_noise=0.02,
    use_flex_spec_target=True, weight_decay=0.013, kendall_activation=True, epoch_stop_smooth=1500,
)
N_ROWS = 400   # -> train 280 (batches 200 + 80), val 60, test 60


class Recorder:
    def __init__(self):
        self.events = []
        self.on = True

    def log(self, kind, value):
        if self.on:
            self.events.append((kind, value))


def net_to_oracle(enc, dec, dis):
    """nn.Module -> oracle state dict (float64 numpy copies)."""
    def lin_stack(seq):
        W, b, a, rm, rv, nbt = [], [], [], [], [], 0
        for m in seq:
            if isinstance(m, torch.nn.Linear):
                W.append(m.weight.detach().numpy().copy())
                b.append(m.bias.detach().numpy().copy())
            elif isinstance(m, torch.nn.PReLU):
                a.append(m.weight.detach().numpy().copy())
            elif isinstance(m, torch.nn.BatchNorm1d):
                rm.append(m.running_mean.numpy().copy())
                rv.append(m.running_var.numpy().copy())
                nbt = int(m.num_batches_tracked)
        return W, b, a, rm, rv, nbt

    We, be, ae, rme, rve, nbe = lin_stack(enc.main)
    Wd, bd, ad, rmd, rvd, nbd = lin_stack(dec.main)
    Ws, bs, as_, _, _, _ = lin_stack(dis.main)
    return dict(E=dict(W=We, b=be, a=ae, rm=rme, rv=rve, nbt=nbe),
                D=dict(W=Wd, b=bd, a=ad, rm=rmd, rv=rvd, nbt=nbd),
                S=dict(W=Ws, b=bs, a=as_))


def flatten_state(prefix, st, out, dtype=np.float32):
    for net in ("E", "D", "S"):
        for k, v in st[net].items():
            if isinstance(v, list):
                for i, x in enumerate(v):
                    out[f"{prefix}.{net}.{k}{i}"] = np.asarray(x, dtype=dtype)
            else:
                out[f"{prefix}.{net}.{k}"] = np.asarray(v)


def round_to_f32_(t):
    with torch.no_grad():
        t.copy_(t.float().double())


def run_case(name, record_epoch, max_epoch, seed, out_path, full_grads=True):
    ref_trainer = ref_shim.import_reference()
    from sc.utils.parameter import Parameters

    torch.set_default_dtype(torch.float64)
    torch.manual_seed(seed)
    cfg_dict = dict(CONFIG, max_epoch=max_epoch)
    ocfg = O.Config.from_dict(cfg_dict)
    spec, aux = O.synthetic_dataset(N_ROWS, ocfg, seed=seed, dtype=np.float32)
    tmp = tempfile.mkdtemp(prefix="raae_golden_")
    csv = os.path.join(tmp, "synthetic.csv")
    ref_shim.write_csv(csv, spec, aux)

    rec = Recorder()
    rec.on = False

    # --- RNG recording wrappers -----------------------------------------------------------
    real_randn, real_randn_like, real_dropout = torch.randn, torch.randn_like, torch.nn.functional.dropout

    def randn(*a, **kw):
        rg = kw.pop("requires_grad", False)
        t = real_randn(*a, **kw).float().double()
        rec.log("randn", t.numpy().copy())
        return t.requires_grad_(rg)

    def randn_like(x, **kw):
        kw.pop("requires_grad", None)
        t = real_randn_like(x, **kw).float().double()
        rec.log("randn_like", t.numpy().copy())
        return t

    def dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        mask = torch.bernoulli(torch.full_like(x, 1.0 - p))
        rec.log("dropout", mask.numpy().astype(np.uint8))
        return x * mask / (1.0 - p)

    torch.randn, torch.randn_like, torch.nn.functional.dropout = randn, randn_like, dropout

    # adversarial_loss builds its BCE labels with an explicit float32 dtype (functions.py:124,127),
    # which drags the whole loss to float32: promote those labels to float64 (values 0/1 unchanged).
    real_ones, real_zeros = torch.ones, torch.zeros

    def _promote(fn):
        def f(*a, **kw):
            if kw.get("dtype") is torch.float32:
                kw["dtype"] = torch.float64
            return fn(*a, **kw)
        return f

    torch.ones, torch.zeros = _promote(real_ones), _promote(real_zeros)

    # GaussianSmoothing builds its taps with an explicit float32 dtype (model.py:188-191), which
    # conv1d rejects against float64 activations: cast the (float32-valued) taps to the input dtype.
    import sc.utils.functions as ref_functions
    RefGS = ref_functions.GaussianSmoothing

    class GS64(RefGS):
        def forward(self, x):
            self.weight = self.weight.to(x.dtype)
            return super().forward(x)

    ref_functions.GaussianSmoothing = GS64

    # --- loss-value recording ---------------------------------------------------------------
    originals = {}
    for fn in ("kendall_constraint", "recon_loss", "mutual_info_loss", "smoothness_loss", "adversarial_loss"):
        originals[fn] = getattr(ref_trainer, fn)

        def make(fn_name, orig):
            def wrapped(*a, **kw):
                out = orig(*a, **kw)
                extra = None
                if fn_name == "kendall_constraint":
                    extra = a[0].detach().numpy().copy()      # the descriptors of this batch
                rec.log("loss:" + fn_name, (float(out.detach()), extra))
                return out
            return wrapped
        setattr(ref_trainer, fn, make(fn, originals[fn]))

    try:
        trainer = ref_trainer.Trainer.from_data(
            csv, igpu=0, verbose=False, work_dir=tmp, config_parameters=Parameters(cfg_dict))
        enc, dec, dis = trainer.encoder, trainer.decoder, trainer.discriminator

        # first encoder call of every batch sees x_noisy (trainer.py:113)
        enc_inputs = []
        hook = enc.register_forward_pre_hook(
            lambda m, inp: enc_inputs.append(inp[0].detach().numpy().copy()) if rec.on else None)

        # grads just before every optimizer step; params after
        steps = []
        for oname, opt in trainer.optimizers.items():
            def make_step(oname, opt, orig_step):
                def step(*a, **kw):
                    if rec.on:
                        g = net_grads(enc, dec, dis)
                    r = orig_step(*a, **kw)
                    if rec.on:
                        steps.append((oname, g, net_to_oracle(enc, dec, dis)))
                    return r
                return step
            opt.step = make_step(oname, opt, opt.step)

        snap = {}
        metrics_log = []

        def callback(epoch, metrics):
            metrics_log.append((epoch, list(metrics)))
            if epoch + 1 == max_epoch:
                hook.remove()        # the hook closure must not reach torch.save (trainer.py:310)
            if epoch + 1 == record_epoch:
                start_recording()

        def start_recording():
            for net in (enc, dec, dis):
                for p in net.parameters():
                    round_to_f32_(p.data)
                for b in net.buffers():
                    if b.dtype.is_floating_point:
                        round_to_f32_(b)
            for opt in trainer.optimizers.values():
                for st in opt.state.values():
                    for k in ("exp_avg", "exp_avg_sq"):
                        if k in st:
                            round_to_f32_(st[k])
            snap["state"] = net_to_oracle(enc, dec, dis)
            snap["opt"] = opt_to_oracle(trainer, enc, dec, dis)
            snap["lrs"] = {k: o.param_groups[0]["lr"] for k, o in trainer.optimizers.items()}
            rec.on = True

        if record_epoch == 0:
            start_recording()
        final_metrics = trainer.train(callback=callback)
        rec.on = False
        end_state = net_to_oracle(enc, dec, dis)
        end_opt = opt_to_oracle(trainer, enc, dec, dis)
    finally:
        torch.randn, torch.randn_like, torch.nn.functional.dropout = real_randn, real_randn_like, real_dropout
        for fn, orig in originals.items():
            setattr(ref_trainer, fn, orig)
        ref_functions.GaussianSmoothing = RefGS
        torch.ones, torch.zeros = real_ones, real_zeros
        torch.set_default_dtype(torch.float32)

    # --- parse the event log in the order of SURVEY.md Appendix E.3 --------------------------
    L1 = CONFIG["n_layers"] - 1
    nS = CONFIG["FC_discriminator_layers"] - 1
    ev = list(rec.events)
    pos = [0]

    def take(kind):
        k, v = ev[pos[0]]
        assert k == kind, f"event {pos[0]}: expected {kind}, got {k}"
        pos[0] += 1
        return v

    n_train = int(N_ROWS * 0.7)
    bs = CONFIG["batch_size"]
    n_batches = (n_train + bs - 1) // bs
    out = {}
    out["config_keys"] = np.array(list(cfg_dict.keys()))
    out["config_vals"] = np.array([repr(v) for v in cfg_dict.values()])
    out["record_epoch"] = np.int64(record_epoch)
    out["spec"] = spec
    out["aux"] = aux
    flatten_state("state0", snap["state"], out)
    flatten_opt("opt0", snap["opt"], out)
    for k, v in snap["lrs"].items():
        out[f"lr0.{k}"] = np.float64(v)
    packed = lambda m: np.packbits(m.astype(np.uint8), axis=None)
    enc_first = [x for x in enc_inputs]
    step_i = 0
    enc_call = 0
    for b in range(n_batches):
        pre = f"b{b}"
        eps_x = take("randn_like")
        B = eps_x.shape[0]
        out[f"{pre}.eps_x"] = eps_x.astype(np.float32)
        x_noisy = enc_first[enc_call]
        # x_noisy is NOT stored: spec[idx] + eps_x * spec_noise reproduces it to 1 ulp (asserted below)
        # recover the batch row indices by matching x_noisy - eps*noise against the train rows
        x_clean = x_noisy - eps_x * CONFIG["spec_noise"]
        idx = np.array([int(np.argmin(np.abs(spec[:n_train].astype(np.float64) - r).sum(axis=1))) for r in x_clean])
        assert np.abs(spec[idx].astype(np.float64) - x_clean).max() < 1e-12
        out[f"{pre}.idx"] = idx.astype(np.int64)
        assert np.abs(spec[idx].astype(np.float64) + eps_x * CONFIG["spec_noise"] - x_noisy).max() < 1e-15
        masks = {}
        masks["E0"] = [take("dropout") for _ in range(L1)]
        masks["D0"] = [take("dropout") for _ in range(L1)]
        out[f"{pre}.z_real"] = take("randn").astype(np.float32)
        out[f"{pre}.S_real_eps"] = take("randn_like").astype(np.float32)
        masks["S_real"] = [take("dropout") for _ in range(nS)]
        out[f"{pre}.S_fake_eps"] = take("randn_like").astype(np.float32)
        masks["S_fake"] = [take("dropout") for _ in range(nS)]
        l_adv = take("loss:adversarial_loss")[0]
        masks["E1"] = [take("dropout") for _ in range(L1)]
        l_aux, desc = take("loss:kendall_constraint")
        assert np.abs(desc - aux[idx]).max() < 1e-12
        masks["E2"] = [take("dropout") for _ in range(L1)]
        masks["D1"] = [take("dropout") for _ in range(L1)]
        l_rec = take("loss:recon_loss")[0]
        masks["E3"] = [take("dropout") for _ in range(L1)]
        out[f"{pre}.z_sample"] = take("randn").astype(np.float32)
        masks["D2"] = [take("dropout") for _ in range(L1)]
        masks["E4"] = [take("dropout") for _ in range(L1)]
        l_mi = take("loss:mutual_info_loss")[0]
        masks["E5"] = [take("dropout") for _ in range(L1)]
        masks["D3"] = [take("dropout") for _ in range(L1)]
        l_sm = take("loss:smoothness_loss")[0]
        for k, ml in masks.items():
            for i, m in enumerate(ml):
                out[f"{pre}.mask.{k}.{i}"] = packed(m)
                out[f"{pre}.maskshape.{k}.{i}"] = np.array(m.shape)
        out[f"{pre}.losses"] = np.array([l_adv, l_aux, l_rec, l_mi, l_sm], dtype=np.float64)
        for ph, oname in zip(O.PHASES, ("adversarial", "correlation", "reconstruction", "mutual_info", "smoothness")):
            sname, g, post = steps[step_i]
            assert sname == oname, (sname, oname)
            step_i += 1
            for net, gd in g.items():
                for k, lst in gd.items():
                    for i, x in enumerate(lst):
                        if x is not None:
                            if full_grads and b == n_batches - 1:   # full tensors for the ragged last batch only
                                out[f"{pre}.grad.{ph}.{net}.{k}{i}"] = x.astype(np.float32)
                            out[f"{pre}.gradsum.{ph}.{net}.{k}{i}"] = np.array([x.sum(), np.sqrt((x ** 2).sum())])
            if ph == "smoothness" and b == 0:
                flatten_state(f"{pre}.post", post, out)
            if os.environ.get("RAAE_GOLDEN_DEBUG"):      # float64 dumps for debugging the oracle
                flatten_state(f"{pre}.dbgpost.{ph}", post, out, dtype=np.float64)
                for net, gd in g.items():
                    for k, lst in gd.items():
                        for i, x in enumerate(lst):
                            if x is not None:
                                out[f"{pre}.dbggrad.{ph}.{net}.{k}{i}"] = x
        enc_call += 6      # six encoder forwards per batch
    # validation block (trainer.py:223-253): recon, kendall, smooth, MI (draws z_sample), adversarial (draws z_real)
    v_rec = take("loss:recon_loss")[0]
    v_aux = take("loss:kendall_constraint")[0]
    v_sm = take("loss:smoothness_loss")[0]
    out["val.z_sample"] = take("randn").astype(np.float32)
    v_mi = take("loss:mutual_info_loss")[0]
    out["val.z_real"] = take("randn").astype(np.float32)
    v_adv = take("loss:adversarial_loss")[0]
    assert pos[0] == len(ev), (pos[0], len(ev))
    # stored in O.PHASES order: adversarial, correlation, reconstruction, mutual_info, smoothness
    out["val.losses"] = np.array([v_adv, v_aux, v_rec, v_mi, v_sm])
    out["val.metrics"] = np.array(metrics_log[-1][1], dtype=np.float64)
    out["final_metrics"] = np.array(final_metrics, dtype=np.float64)
    flatten_state("state1", end_state, out)
    flatten_opt("opt1_sums", end_opt, out, sums_only=True)
    out["lr1"] = np.array([trainer.optimizers[k].param_groups[0]["lr"] for k in O.PHASES])
    np.savez_compressed(out_path, **out)
    print(f"{name}: wrote {out_path} ({os.path.getsize(out_path) / 1e6:.2f} MB), "
          f"losses b0 {out['b0.losses']}, val metrics {out['val.metrics']}")


def net_grads(enc, dec, dis):
    def stack(seq):
        W, b, a = [], [], []
        for m in seq:
            if isinstance(m, torch.nn.Linear):
                W.append(None if m.weight.grad is None else m.weight.grad.numpy().copy())
                b.append(None if m.bias.grad is None else m.bias.grad.numpy().copy())
            elif isinstance(m, torch.nn.PReLU):
                a.append(None if m.weight.grad is None else m.weight.grad.numpy().copy())
        return dict(W=W, b=b, a=a)
    return dict(E=stack(enc.main), D=stack(dec.main), S=stack(dis.main))


def opt_to_oracle(trainer, enc, dec, dis):
    """torch optimizer state -> {name: {t, lr, m:{net:{W,b,a}}, v:...}} for the 5 stepped optimizers."""
    nets = dict(E=enc, D=dec, S=dis)
    hp = O.Config().optimizer_hparams()
    out = {}
    for name in O.PHASES:
        opt = trainer.optimizers[name]
        o = dict(t=0, lr=opt.param_groups[0]["lr"], m={}, v={})

        def mv(p):
            st = opt.state.get(p, {})
            if "step" in st:
                o["t"] = int(st["step"])
            z = np.zeros(tuple(p.shape))
            return (st["exp_avg"].numpy().copy() if "exp_avg" in st else z,
                    st["exp_avg_sq"].numpy().copy() if "exp_avg_sq" in st else z.copy())

        for net_name in hp[name]["nets"]:
            m = dict(W=[], b=[], a=[])
            v = dict(W=[], b=[], a=[])
            for mod in nets[net_name].main:
                if isinstance(mod, torch.nn.Linear):
                    for key, p in (("W", mod.weight), ("b", mod.bias)):
                        a_, b_ = mv(p)
                        m[key].append(a_)
                        v[key].append(b_)
                elif isinstance(mod, torch.nn.PReLU):
                    a_, b_ = mv(mod.weight)
                    m["a"].append(a_)
                    v["a"].append(b_)
            o["m"][net_name], o["v"][net_name] = m, v
        out[name] = o
    return out


def flatten_opt(prefix, opt, out, sums_only=False):
    for name, o in opt.items():
        out[f"{prefix}.{name}.t"] = np.int64(o["t"])
        for mv in ("m", "v"):
            for net, d in o[mv].items():
                for k, lst in d.items():
                    for i, x in enumerate(lst):
                        if sums_only:
                            out[f"{prefix}.{name}.{mv}.{net}.{k}{i}"] = np.array([x.sum(), np.sqrt((x ** 2).sum())])
                        else:
                            out[f"{prefix}.{name}.{mv}.{net}.{k}{i}"] = x.astype(np.float32)


if __name__ == "__main__":
    gdir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    if os.environ.get("RAAE_GOLDEN_DEBUG"):
        gdir = os.environ["RAAE_GOLDEN_DEBUG"]
    os.makedirs(gdir, exist_ok=True)
    run_case("fresh", record_epoch=0, max_epoch=1, seed=11, out_path=os.path.join(gdir, "step_fresh.npz"))
    run_case("warm", record_epoch=3, max_epoch=4, seed=12, out_path=os.path.join(gdir, "step_warm.npz"),
             full_grads=False)
