"""`Trainer` with the reference's surface (`sc/clustering/trainer.py:33-474`): `Trainer.from_data(...)`
builds the networks through `AE_CLS_DICT` and loads the CSV, `Trainer.train(callback)` returns the
5-vector of metrics and writes `final.pt` / `losses.csv` — but the loop body, the validation block,
the metrics and the LR schedulers all run inside the fused sm_100a kernels (rankaae_b200.engine).
"""
import logging
import os
import shutil

import numpy as np
import torch

from .dataloader import get_datasets
from .engine import Engine, auto_ctas_per_trial
from .model import DiscriminatorFC
from .parameter import AE_CLS_DICT, OPTIM_DICT, Parameters

LOSS_HEADER = ("Epoch,Train_D,Val_D,Train_G,Val_G,Train_Aux,Val_Aux,Train_Recon,"
               "Val_Recon,Train_Smooth,Val_Smooth,Train_Mutual_Info,Val_Mutual_Info")     # trainer.py:84-87


def build_modules(p, seed=None):
    """Network construction exactly as Trainer.from_data does it (trainer.py:442-463)."""
    g = p.get if hasattr(p, "get") else (lambda k, d=None: getattr(p, k, d))
    ae_form = g("ae_form", "FC")
    if ae_form not in AE_CLS_DICT:
        raise NotImplementedError(f"ae_form {ae_form!r} is not implemented by the fused path (FC only)")
    if g("use_cnn_discriminator", False):
        raise NotImplementedError("use_cnn_discriminator: true is not implemented by the fused path")
    ctx = torch.random.fork_rng(devices=[]) if seed is not None else None
    if ctx is not None:
        ctx.__enter__()
        torch.manual_seed(int(seed))
    try:
        encoder = AE_CLS_DICT[ae_form]["encoder"](
            nstyle=g("nstyle", 5), dropout_rate=g("dropout_rate", 0.2), dim_in=g("dim_in", 256), n_layers=g("n_layers", 3))
        decoder = AE_CLS_DICT[ae_form]["decoder"](
            nstyle=g("nstyle", 5), dropout_rate=g("dropout_rate", 0.2), last_layer_activation=g("decoder_activation", "ReLu"),
            dim_out=g("dim_out", 256), n_layers=g("n_layers", 3))
        discriminator = DiscriminatorFC(
            nstyle=g("nstyle", 5), dropout_rate=g("dis_dropout_rate", 0.2), noise=g("dis_noise", 0.1),
            layers=g("FC_discriminator_layers", 3))
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None)
    return encoder, decoder, discriminator


REFERENCE_CLASSES = {"FCEncoder": "FCEncoder", "FCDecoder": "FCDecoder", "DiscriminatorFC": "DiscriminatorFC"}


def as_reference_modules(modules):
    """The reference's own `sc.clustering.model` classes carrying the same parameters / buffers, or None when the
    reference package is not importable.  `sc/report/analysis.py:115-121` unpickles final.pt with only the reference on
    the path, so the pickle must name `sc.clustering.model.*` classes."""
    try:
        import sc.clustering.model as ref_model
    except Exception:
        return None
    out = []
    for m in modules:
        cls = getattr(ref_model, REFERENCE_CLASSES[type(m).__name__])
        r = cls(**m._ctor)
        r.load_state_dict(m.state_dict())
        out.append(r.eval())
    return tuple(out)


def save_final(modules, path, reference_classes=None):
    """Writes {"Encoder", "Decoder", "Style Discriminator"} as the reference does (trainer.py:281-283, 310), in eval mode
    (what a consumer that forgets .eval() expects from a finished model).  With the reference importable (or
    `reference_classes=True`) the pickled objects are the reference's own classes, so `sc_generate_report` loads the file
    unchanged; otherwise they are rankaae_b200.model classes and `tools/convert_final_pt.py` converts the file later."""
    enc, dec, dis = (m.eval() for m in modules)
    ref = as_reference_modules((enc, dec, dis)) if reference_classes in (None, True) else None
    if ref is None and reference_classes is True:
        raise RuntimeError("the reference package `sc` is not importable: cannot pickle its classes")
    enc, dec, dis = ref if ref is not None else (enc, dec, dis)
    torch.save({"Encoder": enc, "Decoder": dec, "Style Discriminator": dis}, path)
    return ref is not None


def init_trial_state(engine, trial, cfg, seed=None):
    """Fresh PyTorch-default initialisation of trial `trial` (what every reference engine process does
    on its own, train_sc.py:82-90)."""
    enc, dec, dis = build_modules(cfg, seed=seed)
    engine.load_modules(trial, enc, dec, dis)
    return enc, dec, dis


class Trainer:

    metric_weights = [1.0, -1.0, -0.01, -1.0, -1.0]      # applied in-kernel (aae_kernels.cuh)
    gau_kernel_size = 17

    def __init__(self, encoder, decoder, discriminator, device, train_data, val_data,
                 verbose=True, work_dir='.', tb_logdir="runs", config_parameters=Parameters({}),
                 logger=logging.getLogger("training"), loss_logger=logging.getLogger("losses"),
                 seed=0, epochs_per_call=25):
        self.logger = logger
        self.loss_logger = loss_logger
        self.device = device
        self.encoder, self.decoder, self.discriminator = encoder, decoder, discriminator
        self.train_data, self.val_data = train_data, val_data
        self.verbose = verbose
        self.work_dir = work_dir
        self.tb_logdir = tb_logdir
        self.epoch_stop_smooth = 500                       # trainer.py:59
        self.config_parameters = config_parameters
        self.__dict__.update(config_parameters.to_dict())  # trainer.py:60
        if self.optimizer_name not in OPTIM_DICT:
            raise NotImplementedError(f"optimizer_name {self.optimizer_name!r} is not implemented (AdamW only)")
        self.epochs_per_call = epochs_per_call
        cfg = dict(config_parameters.to_dict())
        cfg.setdefault("epoch_stop_smooth", self.epoch_stop_smooth)
        cfg["ctas_per_trial"] = auto_ctas_per_trial(cfg, 1, device)      # one trial: a thread-block cluster (8 CTAs at batch 1024)
        n_val = val_data[0].shape[0]
        self.engine = Engine(cfg, n_trials=1, device=device, max_rows=max(int(self.batch_size), n_val), seeds=[seed])
        self.engine.load_modules(0, encoder, decoder, discriminator)
        self.engine.bind_dataset(train_data[0], train_data[1], val_data[0], val_data[1])

    def train(self, callback=None):
        if self.verbose:
            self.logger.info(torch.__config__.parallel_info())
        best_combined_metric = 10.0                         # trainer.py:76
        chkpt_dir = f"{self.work_dir}/checkpoints"
        os.makedirs(chkpt_dir, exist_ok=True)
        best_chpt_file = None
        metrics = None
        self.loss_logger.info(LOSS_HEADER)
        chunk = 1 if callback is not None else self.epochs_per_call
        epoch = 0
        while epoch < self.max_epoch:
            n = min(chunk, self.max_epoch - epoch)
            losses, mets = self.engine.train_epochs(epoch, n)
            torch.cuda.synchronize(self.engine.device)
            losses, mets = losses[:, 0].cpu().numpy(), mets[:, 0].cpu().numpy()
            for i in range(n):
                e = epoch + i
                if e % 10 == 0:                             # trainer.py:270-279
                    self.loss_logger.info(f"{e:d},\t" + "".join(f"{v:.6f},\t" for v in losses[i]))
                metrics = [float(v) for v in mets[i, :5]]
                combined_metric = float(mets[i, 5])
                if combined_metric > best_combined_metric:  # trainer.py:298-301
                    # the checkpoint holds the state at the END of this chunk of epochs (the fused kernels run a chunk per
                    # call; with a callback the chunk is one epoch and the file is the epoch's own state, as in the reference):
                    # the file is named after the epoch it was taken at
                    best_combined_metric = combined_metric
                    best_chpt_file = f"{chkpt_dir}/epoch_{epoch + n - 1:06d}_loss_{combined_metric:07.6g}.pt"
                    save_final(self._modules(), best_chpt_file)
                if callback is not None:
                    callback(e, metrics)
            epoch += n
        save_final(self._modules(), f'{self.work_dir}/final.pt')          # trainer.py:310
        if best_chpt_file is not None:
            shutil.copy2(best_chpt_file, f'{self.work_dir}/best.pt')
        return metrics

    def _modules(self):
        self.engine.store_modules(0, self.encoder, self.decoder, self.discriminator)
        return self.encoder, self.decoder, self.discriminator

    def _model_dict(self):
        enc, dec, dis = self._modules()
        return {"Encoder": enc, "Decoder": dec, "Style Discriminator": dis}

    @classmethod
    def from_data(cls, csv_fn, igpu=0, verbose=True, work_dir='.', train_ratio=0.7, validation_ratio=0.15,
                  test_ratio=0.15, config_parameters=Parameters({}), logger=logging.getLogger("from_data"),
                  loss_logger=logging.getLogger("losses"), seed=0):
        p = config_parameters
        assert p.ae_form in AE_CLS_DICT
        ds_train, ds_val, _ = get_datasets(csv_fn, (train_ratio, validation_ratio, test_ratio), n_aux=p.n_aux)
        if not torch.cuda.is_available():
            raise RuntimeError("rankaae_b200 has no CPU path: a CUDA (sm_100a) device is required")
        if verbose:
            logger.info("Use GPU")
        device = torch.device(f"cuda:{igpu}")
        encoder, decoder, discriminator = build_modules(p, seed=seed)
        return cls(encoder, decoder, discriminator, device, ds_train.tensors(), ds_val.tensors(),
                   verbose=verbose, work_dir=work_dir, config_parameters=p, logger=logger, loss_logger=loss_logger,
                   seed=seed)
