"""CPU: the C-ABI library loads, exports every symbol include/deepsee_b200.h declares, and fails
loudly (no CPU fallback) when asked to compute without a B200."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "deepsee_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsee_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from deepsee_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libdeepsee_b200.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "deepsee_b200/_lib.py SYMBOLS is out of sync with the header"
    assert lib.dsee_version() == _lib.ABI_VERSION


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    from deepsee_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    rc = lib.dsee_bn_eval_affine(ctypes.addressof(buf), ctypes.addressof(buf), 1e-5, 16,
                                 ctypes.addressof(buf), ctypes.addressof(buf), None)
    assert rc != 0
    assert b"no CPU fallback" in lib.dsee_last_error() or b"sm_100" in lib.dsee_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_ops_reject_cpu_tensors():
    from deepsee_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.bn_eval_affine(torch.zeros(8), torch.ones(8), 1e-5)


def test_argument_validation_message():
    from deepsee_b200 import _lib
    lib = _lib.load()
    rc = lib.dsee_resize_labels(None, None, 1, 4, 4, 2, 2, None)
    assert rc == -1 and b"bad argument" in lib.dsee_last_error()
