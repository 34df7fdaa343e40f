"""Load the UNMODIFIED reference implementation of the hot path, when it is present.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  ``/root/reference`` exists only in
the build container (never on the GPU box), so everything here is optional:
``available()`` says whether it can be used; tests that need it skip otherwise, and
``tests/golden/make_golden.py`` uses it to produce the committed golden vectors.

``import codebase...`` fails without fairseq (codebase/__init__.py:6 imports every
sub-package, which register fairseq models).  The hot-path files themselves only need
torch, so they are loaded by file path under stub parent packages:

  codebase/utils/functions.py, monotonic_attention.py, p_choose_strategy.py
  codebase/models/torch_cif/cif.py
  codebase/modules/monotonic_multihead_attention.py   (with a minimal stand-in for
      fairseq.modules.MultiheadAttention and the registry decorator)

No reference source is copied; the files are executed where they lie.
"""
import importlib.util
import os
import sys
import types
from typing import Optional

REF_ROOT = os.environ.get("SIMULST_REFERENCE_ROOT", "/root/reference")

_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "codebase/utils/monotonic_attention.py"))


def _stub_package(name: str):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        mod.__path__ = []  # mark as package
        sys.modules[name] = mod
    return sys.modules[name]


def _load(mod_name: str, rel_path: str):
    if mod_name in _cache:
        return _cache[mod_name]
    path = os.path.join(REF_ROOT, rel_path)
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    spec.loader.exec_module(mod)
    _cache[mod_name] = mod
    return mod


def load_utils():
    """Returns (functions, monotonic_attention, p_choose_strategy) reference modules."""
    _stub_package("codebase")
    _stub_package("codebase.utils")
    fn = _load("codebase.utils.functions", "codebase/utils/functions.py")
    ma = _load("codebase.utils.monotonic_attention", "codebase/utils/monotonic_attention.py")
    pc = _load("codebase.utils.p_choose_strategy", "codebase/utils/p_choose_strategy.py")
    return fn, ma, pc


def load_cif():
    """Returns the reference torch_cif.cif module."""
    return _load("_simulst_ref_cif", "codebase/models/torch_cif/cif.py")


def _install_fairseq_stub():
    """Smallest stand-in for what monotonic_multihead_attention.py needs from fairseq:
    a MultiheadAttention base with q/k/v/out projections and incremental-state accessors."""
    if "fairseq.modules" in sys.modules and hasattr(sys.modules["fairseq.modules"], "_simulst_stub"):
        return
    import torch.nn as nn

    class MultiheadAttention(nn.Module):
        def __init__(self, embed_dim, num_heads, kdim=None, vdim=None, dropout=0.0,
                     bias=True, encoder_decoder_attention=False, **kw):
            super().__init__()
            self.embed_dim = embed_dim
            self.kdim = kdim if kdim is not None else embed_dim
            self.vdim = vdim if vdim is not None else embed_dim
            self.qkv_same_dim = self.kdim == embed_dim and self.vdim == embed_dim
            self.num_heads = num_heads
            self.head_dim = embed_dim // num_heads
            self.scaling = self.head_dim ** -0.5
            self.k_proj = nn.Linear(self.kdim, embed_dim, bias=bias)
            self.v_proj = nn.Linear(self.vdim, embed_dim, bias=bias)
            self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
            self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)

        def get_incremental_state(self, incremental_state, key):
            if incremental_state is None:
                return None
            return incremental_state.get((id(self), key))

        def set_incremental_state(self, incremental_state, key, value):
            if incremental_state is not None:
                incremental_state[(id(self), key)] = value
            return incremental_state

    fairseq = _stub_package("fairseq")
    fmods = _stub_package("fairseq.modules")
    fmods.MultiheadAttention = MultiheadAttention
    fmods._simulst_stub = True
    fairseq.modules = fmods


def load_mma_module():
    """Returns the reference modules.monotonic_multihead_attention module (classes
    MonotonicAttention, MonotonicInfiniteLookbackAttention, ...)."""
    load_utils()
    _install_fairseq_stub()
    pkg = _stub_package("codebase.modules")
    if not hasattr(pkg, "register_monotonic_attention"):
        def register_monotonic_attention(name):
            def deco(cls):
                return cls
            return deco
        pkg.register_monotonic_attention = register_monotonic_attention
    return _load("codebase.modules.monotonic_multihead_attention",
                 "codebase/modules/monotonic_multihead_attention.py")


def make_attention(kind: str = "infinite_lookback", embed_dim: int = 64, heads: int = 4,
                   mass_preservation: bool = True, eps: float = 1e-6,
                   energy_bias: bool = True, seed: Optional[int] = 0):
    """Instantiate a reference attention module on CPU with deterministic weights."""
    import argparse
    import torch
    mod = load_mma_module()
    cls = {"hard_aligned": mod.MonotonicAttention,
           "infinite_lookback": mod.MonotonicInfiniteLookbackAttention}[kind]
    args = argparse.Namespace(
        decoder_embed_dim=embed_dim, decoder_attention_heads=heads,
        encoder_embed_dim=embed_dim, attention_dropout=0.0,
        attention_eps=eps, mass_preservation=mass_preservation,
        noise_mean=0.0, noise_var=1.0, energy_bias_init=-2.0, energy_bias=energy_bias)
    if seed is not None:
        torch.manual_seed(seed)
    return cls(args)
