"""MMA training step for callers whose activations live in HOST memory.

`MMAHostPipeline.step()` is the host-buffer entry of the fused path
(monotonic_attention_process_train steps 2-3, reference
codebase/modules/monotonic_multihead_attention.py:318-347, plus its backward): p_choose and
soft_energy are read from pinned host tensors, grad_p / grad_energy are written to pinned host
tensors.  Rows ((batch, head) pairs) never interact, so the batch is cut into row
chunks that flow through three CUDA streams

    copy-in stream :  H2D p[k], e[k]
    compute streams:  simulst_mma_train_fwd(k) ; simulst_mma_train_bwd(k)      (round-robin)
    copy-out stream:  D2H grad_p[k], grad_e[k]

so the upload of chunk k+1, the kernels of chunk k and the download of chunk k-1 overlap and
PCIe runs in both directions at once.  Everything is ordered after the caller's current stream
at entry and joined back into it at exit: a call behaves like one (long) operation on that
stream.  Device buffers are allocated once in the constructor; the library itself neither
allocates nor synchronises.
"""
from typing import Optional

import torch
from torch import Tensor

from . import _lib


class MMAHostPipeline:
    def __init__(self, n_rows: int, tgt_len: int, src_len: int, dtype: torch.dtype = torch.bfloat16,
                 device=None, chunks: int = 8, compute_streams: int = 2, eps: float = 1e-6,
                 mass_preservation: bool = True, soft: bool = True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.BackendUnavailable("MMAHostPipeline needs a CUDA device (no CPU fallback)")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n, self.t, self.s = int(n_rows), int(tgt_len), int(src_len)
        self.dtype, self.eps, self.soft = dtype, float(eps), bool(soft)
        self.flags = (_lib.MMA_MASS_PRESERVATION if mass_preservation else 0) | (_lib.MMA_SOFT if soft else 0)
        self.dt_enum = _lib.dtype_enum(dtype)
        chunks = max(1, min(int(chunks), self.n))
        base, rem = divmod(self.n, chunks)
        self.bounds = []
        lo = 0
        for k in range(chunks):
            hi = lo + base + (1 if k < rem else 0)
            self.bounds.append((lo, hi))
            lo = hi
        shape = (self.n, self.t, self.s)
        d = self.dev
        self.p = torch.empty(shape, dtype=dtype, device=d)
        self.e = torch.empty(shape, dtype=dtype, device=d) if soft else None
        self.alpha = torch.empty(shape, dtype=torch.float32, device=d)
        self.beta = torch.empty(shape, dtype=torch.float32, device=d) if soft else None
        self.side = torch.empty((self.n, self.t, 2), dtype=torch.float32, device=d)
        self.grad_p = torch.empty(shape, dtype=dtype, device=d)
        self.grad_e = torch.empty(shape, dtype=dtype, device=d) if soft else None
        self.status = _lib.status_word(d)
        self.s_in = torch.cuda.Stream(d)
        self.s_out = torch.cuda.Stream(d)
        self.s_comp = [torch.cuda.Stream(d) for _ in range(max(1, int(compute_streams)))]
        self.ev_in = [torch.cuda.Event() for _ in self.bounds]
        self.ev_comp = [torch.cuda.Event() for _ in self.bounds]

    def _check_host(self, t: Optional[Tensor], what: str):
        if t is None:
            raise ValueError(f"{what} is required")
        if t.is_cuda or tuple(t.shape) != (self.n, self.t, self.s) or t.dtype != self.dtype or not t.is_contiguous():
            raise ValueError(f"{what} must be a contiguous host tensor [{self.n},{self.t},{self.s}] of {self.dtype}")
        if not t.is_pinned():
            raise ValueError(f"{what} must be pinned (page-locked) for asynchronous copies")

    def step(self, p_host: Tensor, e_host: Optional[Tensor], grad_alpha: Optional[Tensor],
             grad_beta: Optional[Tensor], grad_p_host: Tensor, grad_e_host: Optional[Tensor]):
        """One forward + backward over host-resident activations.  `grad_alpha` / `grad_beta`
        are the upstream gradients (device, fp32, [N,T,S]; produced by the loss on the GPU).
        Returns (alpha, beta): device tensors owned by the pipeline, valid until the next step."""
        self._check_host(p_host, "p_host")
        self._check_host(grad_p_host, "grad_p_host")
        if self.soft:
            self._check_host(e_host, "e_host")
            self._check_host(grad_e_host, "grad_e_host")
        for g, what in ((grad_alpha, "grad_alpha"), (grad_beta, "grad_beta")):
            if g is not None and (not g.is_cuda or g.dtype != torch.float32 or not g.is_contiguous()
                                  or tuple(g.shape) != (self.n, self.t, self.s)):
                raise ValueError(f"{what} must be a contiguous fp32 CUDA tensor [{self.n},{self.t},{self.s}]")
        lib, DT, T, S = self.lib, self.dt_enum, self.t, self.s
        cur = torch.cuda.current_stream(self.dev)
        self.s_in.wait_stream(cur)
        self.s_out.wait_stream(cur)
        for sc in self.s_comp:
            sc.wait_stream(cur)
        with torch.cuda.device(self.dev):
            for k, (lo, hi) in enumerate(self.bounds):
                rows = hi - lo
                with torch.cuda.stream(self.s_in):
                    self.p[lo:hi].copy_(p_host[lo:hi], non_blocking=True)
                    if self.soft:
                        self.e[lo:hi].copy_(e_host[lo:hi], non_blocking=True)
                    self.ev_in[k].record(self.s_in)
                sc = self.s_comp[k % len(self.s_comp)]
                sc.wait_event(self.ev_in[k])
                st = sc.cuda_stream
                rc = lib.simulst_mma_train_fwd(
                    self.p[lo:hi].data_ptr(), DT, self.e[lo:hi].data_ptr() if self.soft else None, DT, None,
                    self.alpha[lo:hi].data_ptr(), self.beta[lo:hi].data_ptr() if self.soft else None,
                    self.side[lo:hi].data_ptr(), rows, T, S, self.eps, 0, self.flags,
                    self.status.data_ptr(), st)
                _lib.check(rc, "simulst_mma_train_fwd")
                rc = lib.simulst_mma_train_bwd(
                    self.p[lo:hi].data_ptr(), DT, self.e[lo:hi].data_ptr() if self.soft else None, DT, None,
                    self.alpha[lo:hi].data_ptr(), self.side[lo:hi].data_ptr(),
                    grad_alpha[lo:hi].data_ptr() if grad_alpha is not None else None,
                    grad_beta[lo:hi].data_ptr() if (self.soft and grad_beta is not None) else None,
                    self.grad_p[lo:hi].data_ptr(), DT,
                    self.grad_e[lo:hi].data_ptr() if self.soft else None, DT,
                    rows, T, S, self.eps, 0, self.flags, st)
                _lib.check(rc, "simulst_mma_train_bwd")
                self.ev_comp[k].record(sc)
                self.s_out.wait_event(self.ev_comp[k])
                with torch.cuda.stream(self.s_out):
                    grad_p_host[lo:hi].copy_(self.grad_p[lo:hi], non_blocking=True)
                    if self.soft:
                        grad_e_host[lo:hi].copy_(self.grad_e[lo:hi], non_blocking=True)
        cur.wait_stream(self.s_out)
        for sc in self.s_comp:
            cur.wait_stream(sc)
        _lib.maybe_check(self.dev)
        return self.alpha, (self.beta if self.soft else self.alpha)

    @property
    def launches_per_step(self) -> int:
        return 2 * len(self.bounds)

    @property
    def h2d_bytes_per_step(self) -> int:
        per = self.n * self.t * self.s * self.p.element_size()
        return per * (2 if self.soft else 1)

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.h2d_bytes_per_step
