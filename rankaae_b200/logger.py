"""File/stream logger factory with the reference's formats (`sc/utils/logger.py:5-34`): `messages.txt`
lines are '%m-%d %H:%M LEVEL:  msg', `losses.csv` lines are the bare message."""
import logging
import os


def create_logger(logger_name, log_path=None, append=False, simple_fmt=False):
    if log_path is not None and not append and os.path.isfile(log_path):
        open(log_path, "w").close()
    logger = logging.getLogger(logger_name)
    logger.setLevel(logging.DEBUG)
    for h in list(logger.handlers):
        logger.removeHandler(h)
    handler = logging.StreamHandler() if log_path is None else logging.FileHandler(log_path)
    handler.setLevel(logging.DEBUG)
    if simple_fmt:
        handler.setFormatter(logging.Formatter(fmt="%(message)s"))
    else:
        handler.setFormatter(logging.Formatter(fmt='%(asctime)s %(levelname)s:  %(message)s', datefmt='%m-%d %H:%M'))
    logger.addHandler(handler)
    logger.propagate = False
    return logger
