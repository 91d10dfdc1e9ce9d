"""End-of-training band of the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Runs `/root/reference`'s own `Trainer.from_data(...).train()` (float32, CPU, exactly as shipped apart from the import
shims in oracle/ref_shim.py) for several seeds on a small synthetic CSV and records, per seed, the trainer's metric
5-vector (trainer.py:294-295) and the per-descriptor Spearman correlation between the validation latents and the
descriptors (the definition of sc/report/analysis.py:374-376).  The result is committed as
tests/golden/e2e_band_ref.json; tests/test_e2e_band_gpu.py trains the same configuration with the fused path and
requires its seed-averaged metrics to lie inside the reference band.

    python oracle/make_e2e_band.py                            # 60 epochs, ~5 min on 8 cores (24 seeds)
    RAAE_BAND_EPOCHS=300 python oracle/make_e2e_band.py       # the longer budget -> e2e_band_ref_long.json (~25 min)
"""
import json
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

N_ROWS, SEEDS = 1400, list(range(24))
MAX_EPOCH = int(os.environ.get("RAAE_BAND_EPOCHS", "60"))          # 60: e2e_band_ref.json; 300: e2e_band_ref_long.json
CONFIG = dict(
    data_file="synthetic.csv", trials=1, timeout=10, verbose=False, max_epoch=MAX_EPOCH, batch_size=256,
    gradient_reversal=True, alpha_flat_step=739, alpha_limit=0.7172, decoder_activation="Softplus",
    dis_beta=1.1, dis_dropout_rate=0.056, dis_noise=0.56, gen_beta=1.1,
    n_aux=5, nstyle=6, ae_form="FC", dim_in=256, dim_out=256, n_layers=5, FC_discriminator_layers=3,
    use_cnn_discriminator=False, dropout_rate=0.04, sch_factor=0.1, sch_patience=100,
    lr_base=0.001, lr_ratio_Corr=10, lr_ratio_Mutual=1, lr_ratio_Reconn=10, lr_ratio_Smooth=1,
    lr_ratio_dis=1, lr_ratio_gen=10, optimizer_name="AdamW", spec_noise=0.02,
    use_flex_spec_target=True, weight_decay=0.01, kendall_activation=True, epoch_stop_smooth=1500,
)
DATA_SEED = 123


def one_seed(seed):
    import torch
    from scipy.stats import spearmanr
    from oracle import aae_oracle as O
    from oracle import ref_shim
    torch.set_num_threads(1)
    ref_trainer = ref_shim.import_reference()
    from sc.utils.parameter import Parameters
    spec, aux = O.synthetic_dataset(N_ROWS, O.Config.from_dict(CONFIG), seed=DATA_SEED, dtype=np.float32)
    tmp = tempfile.mkdtemp(prefix="raae_band_")
    csv = os.path.join(tmp, "synthetic.csv")
    ref_shim.write_csv(csv, spec, aux)
    torch.manual_seed(seed)
    trainer = ref_trainer.Trainer.from_data(csv, igpu=0, verbose=False, work_dir=tmp, config_parameters=Parameters(dict(CONFIG)))
    metrics = trainer.train()
    n_train, n_val = int(N_ROWS * 0.7), int(N_ROWS * 0.15)
    enc = trainer.encoder.eval()
    with torch.no_grad():
        z = enc(torch.from_numpy(spec[n_train:n_train + n_val])).numpy()
    rho = [float(spearmanr(z[:, k], aux[n_train:n_train + n_val, k]).correlation) for k in range(CONFIG["n_aux"])]
    return dict(seed=seed, metrics=[float(m) for m in metrics], descriptor_spearman=rho)


if __name__ == "__main__":
    with mp.get_context("spawn").Pool(min(8, os.cpu_count())) as pool:
        res = pool.map(one_seed, SEEDS)
    out = dict(config=CONFIG, n_rows=N_ROWS, data_seed=DATA_SEED, seeds=SEEDS, runs=res,
               note="reference = unmodified sc.clustering.trainer.Trainer, float32 CPU, torch.manual_seed(seed) before from_data")
    path = os.path.join(os.path.dirname(HERE), "tests", "golden",
                        "e2e_band_ref.json" if MAX_EPOCH == 60 else "e2e_band_ref_long.json")
    json.dump(out, open(path, "w"), indent=1)
    m = np.array([r["metrics"] for r in res])
    rho = np.array([r["descriptor_spearman"] for r in res])
    print("metrics mean", m.mean(0), "std", m.std(0))
    print("descriptor spearman mean", rho.mean(0), "std", rho.std(0))
