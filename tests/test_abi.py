"""CPU: the C-ABI library loads and exports every symbol include/simulst_b200.h declares; the
host wrappers refuse to run without CUDA (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "simulst_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(simulst_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from simulst_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.simulst_version() == 100
    assert b"argument" in lib.simulst_error_string(-1)


def test_argument_errors_are_reported_without_a_gpu():
    from simulst_b200 import _lib
    lib = _lib.load()
    assert lib.simulst_mma_set_config(7, 3) == -1
    assert lib.simulst_mma_set_config(0, 0) == 0
    # null pointers / bad dtype are rejected before any CUDA call
    assert lib.simulst_mma_train_fwd(None, 0, None, 0, None, None, None, None, 1, 1, 1, 1e-6, 0, 0, None, None) == -1
    assert lib.simulst_cif_plan(None, 0, None, None, None, None, None, None, None, None, None, 2, 1, 1, 1.0, None, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import simulst_b200
    from simulst_b200.utils.monotonic_attention import expected_alignment_from_p_choose
    from simulst_b200.models.torch_cif import cif_function
    with pytest.raises(simulst_b200.BackendUnavailable):
        expected_alignment_from_p_choose(torch.rand(1, 2, 8))
    with pytest.raises(simulst_b200.BackendUnavailable):
        cif_function(torch.rand(1, 4, 2), torch.rand(1, 4))


def test_multimem_allreduce_rejects_bad_arguments_without_touching_the_device():
    """simulst_multimem_allreduce_f32 validates on the host before any CUDA call: null buffer, rank outside
    the world, a length that is not a multiple of 4 floats, a misaligned multicast address."""
    from simulst_b200 import _lib
    lib = _lib.load()
    assert lib.simulst_multimem_allreduce_f32(None, 1024, 0, 2, 2, None) == -1
    assert lib.simulst_multimem_allreduce_f32(4096, 1024, 2, 2, 2, None) == -1
    assert lib.simulst_multimem_allreduce_f32(4096, 1024, 0, 2, 0, None) == -1
    assert lib.simulst_multimem_allreduce_f32(4096, 1022, 0, 2, 2, None) == -5
    assert lib.simulst_multimem_allreduce_f32(4100, 1024, 0, 2, 2, None) == -5
    assert lib.simulst_multimem_allreduce_f32(4096, 0, 0, 2, 2, None) == 0
    assert lib.simulst_mma_pooled_workspace_bytes(512, 128, 1024, 8) > 2 * 512 * 128 * 128 * 4
    assert lib.simulst_mma_pooled_workspace_bytes(1, 1, 0, 8) < 0


def test_row_pitch_queries_and_argument_checks_without_a_gpu():
    """simulst_mma_out_pitch is a host-side size query; the row-pitch entry points reject a pitch below S and
    the cluster knobs reject shapes outside their tables before any CUDA call."""
    from simulst_b200 import _lib
    lib = _lib.load()
    # dense rows that already qualify keep their pitch; everything else gets 16-byte rows with room for whole threads
    for s, want in ((1024, 1024), (1504, 1504), (8, 8), (4096, 4096), (1500, 1504), (999, 1000), (1001, 1008),
                    (37, 40), (3001, 3008)):
        assert lib.simulst_mma_out_pitch(s) == want, s
    for s in (1, 7, 130, 2047, 4090, 5003, 6100, 6143, 9000, 16384):
        ld = lib.simulst_mma_out_pitch(s)
        assert ld >= s and ld % 8 == 0 and ld - s < 32, (s, ld)
    assert lib.simulst_mma_out_pitch(0) < 0 and lib.simulst_mma_out_pitch(16385) < 0
    dummy = 256     # a non-null, 256-byte aligned "pointer": the calls below return before touching it
    assert lib.simulst_mma_train_fwd_pitched(dummy, 1, 99, dummy, 1, 100, None, dummy, 104, dummy, 104, None, None,
                                             2, 3, 100, 1e-6, 0, 3, None, None) == -2      # ld_p < S
    assert lib.simulst_mma_train_bwd_pitched(dummy, 1, 100, dummy, 1, 100, None, dummy, 104, None, dummy, 100, dummy,
                                             100, None, dummy, 1, 96, dummy, 1, 104, 2, 3, 100, 1e-6, 0, 2, None) == -2
    assert lib.simulst_mma_set_cluster(3) == -1 and lib.simulst_mma_set_cluster(1) == 0
    assert lib.simulst_mma_set_cluster_shape(3, 128) == -1 and lib.simulst_mma_set_cluster_shape(4, 160) == -1
    assert lib.simulst_mma_set_cluster_shape(0, 0) == 0
