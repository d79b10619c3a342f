"""Minimal rank-aware logger with the reference's call names (ppsci/utils/logger.py:61-232)."""
from __future__ import annotations

import logging
import os
import sys

_logger = logging.getLogger("ppsci_b200")
_initialized = False


def init_logger(name: str = "ppsci_b200", log_file=None, log_level: int = logging.INFO):
    global _initialized
    _logger.setLevel(log_level)
    _logger.propagate = False
    if not _initialized:
        h = logging.StreamHandler(sys.stdout)
        h.setFormatter(logging.Formatter("[%(asctime)s] %(name)s %(levelname)s: %(message)s", "%Y/%m/%d %H:%M:%S"))
        _logger.addHandler(h)
        _initialized = True
    if log_file is not None and int(os.environ.get("RANK", "0")) == 0:
        os.makedirs(os.path.dirname(os.path.abspath(log_file)), exist_ok=True)
        fh = logging.FileHandler(log_file, "a")
        fh.setFormatter(logging.Formatter("[%(asctime)s] %(name)s %(levelname)s: %(message)s", "%Y/%m/%d %H:%M:%S"))
        _logger.addHandler(fh)


def _rank0() -> bool:
    return int(os.environ.get("RANK", "0")) == 0


def info(msg, *a):
    if not _initialized:
        init_logger()
    if _rank0():
        _logger.info(msg, *a)


def message(msg, *a):
    info(msg, *a)


def debug(msg, *a):
    if not _initialized:
        init_logger()
    if _rank0():
        _logger.debug(msg, *a)


def warning(msg, *a):
    if not _initialized:
        init_logger()
    if _rank0():
        _logger.warning(msg, *a)


def error(msg, *a):
    if not _initialized:
        init_logger()
    _logger.error(msg, *a)
