"""simulst_b200 -- B200-native (sm_100a) streaming-alignment hot path of simulst.

The package mirrors the reference's module layout for the functions on the path
(``codebase/utils/...`` -> ``simulst_b200/utils/...`` etc.) so a fairseq user-dir can import
them under the same names; the math runs in hand-written CUDA reached through the C ABI in
``include/simulst_b200.h``.  There is no CPU path: operators raise if the extension or a
CUDA device is missing.
"""
from . import _lib
from ._lib import (BackendUnavailable, assume_right_padding, check_status, set_pitched_outputs,  # noqa: F401
                   set_strict)

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH


def launch_count() -> int:
    return int(_lib.load().simulst_launch_count())


def reset_launch_count() -> None:
    _lib.load().simulst_reset_launch_count()
