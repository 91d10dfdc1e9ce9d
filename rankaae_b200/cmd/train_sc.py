#!/usr/bin/env python
"""`train_sc -c fix_config.yaml -w work_dir` with the reference's flags, config keys and output layout
(`sc/cmd/train_sc.py:105-156`).  `trials > 1` no longer needs an ipcluster: all trials of the config train
concurrently on the GPU; launch under torchrun (`--nproc-per-node N`) to spread them over N GPUs."""
import argparse
import os
import time

import numpy as np
import torch

from rankaae_b200.ensemble import gather_results, run_ensemble
from rankaae_b200.logger import create_logger
from rankaae_b200.parameter import Parameters


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-c', '--config', type=str, required=True,
                        help='Config for training parameter in YAML format')
    parser.add_argument('-w', "--work_dir", type=str, default='.',
                        help="Working directory to write the output files")
    args = parser.parse_args(argv)

    work_dir = os.path.abspath(os.path.expanduser(args.work_dir))
    train_config = Parameters.from_yaml(os.path.join(work_dir, args.config))
    assert os.path.exists(work_dir)
    verbose = train_config.get("verbose", False)
    trials = train_config.get("trials", 1)
    data_file = os.path.join(work_dir, train_config.get("data_file", None))
    timeout = train_config.get("timeout", 10)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    logger = create_logger("Main training:", f'{work_dir}/main_process_message.txt', append=True) if rank == 0 else None
    if rank == 0:
        logger.info("START")
        logger.info("Running with {} process(es).".format(world))
    start = time.time()
    local_res = run_ensemble(work_dir, train_config, data_file, trials, verbose=verbose, device=f"cuda:{local}",
                             rank=rank, world=world, timeout_hours=timeout)
    rows = [m + [t] for m, t in local_res]
    res = gather_results(rows, trials, world, rank, device=f"cuda:{local}")
    if rank == 0:
        time_trials = res[:, 5]
        logger.info(f"Time used for each trial: {time_trials.mean():.2f} +/- {time_trials.std():.2f}s.\n" +
                    ' '.join([f"{t:.2f}s" for t in time_trials]))
        end = time.time()
        logger.info(f"Total time used: {end - start:.2f}s for {trials} trails " +
                    f"({(end - start) / trials:.2f} each on average).")
        logger.info("END\n\n")
    if world > 1:
        torch.distributed.destroy_process_group()
    return res


if __name__ == '__main__':
    main()
