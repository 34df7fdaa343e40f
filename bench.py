#!/usr/bin/env python
"""Benchmark of the streaming-alignment hot path (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over one batch: the MMA training shape B=64 utterances x
H=8 heads, tgt 128, src 1024, bf16 p_choose / soft energy in, fp32 alpha / beta out, fused
forward + fused backward (2 kernel launches).  Multi-GPU shards the utterance batch: every rank
processes its own 64 utterances (weak scaling, no data-path collective) and all-reduces a
stand-in parameter-gradient buffer (1.6 M fp32, the MMA decoder's q/k projections) over NCCL.

Prints ONE JSON line (rank 0).  `value` is whole-job elements/s with inputs resident in HBM;
`e2e` is the same metric through the C-ABI with HOST (pinned) buffers, copies inside the timed
region; `roofline` is the dominant kernel (backward) against the measured HBM peak;
`cpu_baseline` is the oracle port (the reference's algorithm, torch CPU ops) on this box's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mma_expected_alignment_fwd_bwd_elements_per_s"
UNIT = "elements/s"          # element = one (batch, head, tgt, src) cell
B, H, T, S = 64, 8, 128, 1024
N_ROWS = B * H
EPS = 1e-6
BYTES_FWD, BYTES_BWD = 12, 20    # algorithmic bytes per element, bf16 in (SURVEY 8d / DESIGN.md)
GRAD_BUF_ELEMS = 6 * 4 * 256 * 256   # stand-in: 6 layers x (q,k,q_soft,k_soft) x 256x256
CPU_SAMPLE_ROWS = 64                 # bounded CPU sample: 64 of the 512 rows, full T x S


def config(n_gpus):
    return {
        "workload": "MMA infinite-lookback training shape, fused expected alignment + mass "
                    "preservation + expected soft attention, forward+backward",
        "B_per_gpu": B, "H": H, "tgt": T, "src": S, "global_batch": B * n_gpus,
        "inputs": "bf16 p_choose + soft_energy", "outputs": "fp32 alpha + beta; bf16 grads",
        "parallelism": f"dp{n_gpus} (utterance batch sharded, NCCL all-reduce of a "
                       f"{GRAD_BUF_ELEMS * 4 / 1e6:.1f} MB grad buffer)" if n_gpus > 1 else "single GPU",
        "l2": "inputs+outputs 2.1 GB per step >> 126 MB L2 (no flush needed)",
    }


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram bytes per launch from the committed ncu summary of the same shape, else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) > 8 for k in range(4)
                          if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU arm
def _reference_functions():
    """(kind, expected_alignment_from_p_choose, mass_preservation, expected_soft_attention): the
    UNMODIFIED reference functions (codebase/utils/monotonic_attention.py, executed from
    baseline/_ref on the GPU box, /root/reference in the build container) -> kind "reference";
    the oracle port only if those files are absent -> kind "port"."""
    from oracle import ref_loader
    if ref_loader.available():
        _, ma, _ = ref_loader.load_utils()
        return "reference", ma.expected_alignment_from_p_choose, ma.mass_preservation, ma.expected_soft_attention
    from oracle import mma as omma
    return "port", omma.expected_alignment_from_p_choose, omma.mass_preservation, omma.expected_soft_attention


def cpu_port_step(p, e, ga, gb):
    """Steps 2-3 of the reference's monotonic_attention_process_train
    (modules/monotonic_multihead_attention.py:318-347) with its own functions + autograd backward."""
    _, align, mass, soft = _reference_functions()
    p = p.detach().requires_grad_()
    e = e.detach().requires_grad_()
    alpha = align(p.float(), None, eps=EPS)
    alpha = mass(alpha, None)
    beta = soft(alpha, e, padding_mask=None, chunk_size=None, eps=EPS)
    ((alpha * ga).sum() + (beta * gb).sum()).backward()
    return alpha.detach(), beta.detach(), p.grad, e.grad


def parity_rows(p_rows, e_rows, ga_rows, gb_rows, got):
    """The timed tensors checked against the reference on the rows given (CPU, fp32 on the
    bf16-rounded inputs).  alpha / beta: rtol 1e-5 + atol 1e-6*scale; bf16 gradients: one bf16
    rounding step (2^-8) + the fp32 accumulation floor.  Raises on failure."""
    import torch
    a_r, b_r, gp_r, ge_r = cpu_port_step(p_rows.float(), e_rows.float(), ga_rows, gb_rows)
    rep = {}
    floor = 4e-7 * S * float(max(ga_rows.abs().max(), gb_rows.abs().max()))
    for name, x, y, rtol, extra in (("alpha", got[0], a_r, 1e-5, 0.0), ("beta", got[1], b_r, 1e-5, 0.0),
                                    ("grad_p", got[2], gp_r, 2.0 ** -8, floor),
                                    ("grad_energy", got[3], ge_r, 2.0 ** -8, floor)):
        x, y = x.double().cpu(), y.double()
        scale = float(y.abs().max())
        err = (x - y).abs()
        allowed = rtol * y.abs() + 1e-6 * scale + extra
        rep[name] = {"max_abs_err": float(err.max()), "scale": scale,
                     "worst_err_over_allowed": float((err / allowed).max())}
        if bool((err > allowed).any()) or bool(torch.isnan(x).any()):
            raise SystemExit(f"bench parity check failed for {name}: {rep[name]}")
    return rep


def cpu_inputs(rows):
    import torch
    g = torch.Generator().manual_seed(1234)
    p = torch.sigmoid(torch.randn(rows, T, S, generator=g) - 2.0)
    e = torch.randn(rows, T, S, generator=g)
    ga = (torch.arange(1, S + 1).float() / S).expand(rows, T, S) + 1e-2 * torch.randn(rows, T, S, generator=g)
    gb = torch.randn(rows, T, S, generator=g)
    return p, e, ga, gb


def time_cpu(reps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    args = cpu_inputs(CPU_SAMPLE_ROWS)
    for _ in range(warmup):
        cpu_port_step(*args)
    best = float("inf")
    total = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_port_step(*args)
        dt = time.perf_counter() - t0
        best = min(best, dt)
        total += dt
    elems = CPU_SAMPLE_ROWS * T * S
    return elems / best, elems / (total / reps), total / reps, torch.get_num_threads()


def time_gpu_eager(dev):
    """SURVEY 8d "reference GPU path" row: the same restatement of the reference's primitive
    sequence (oracle port; the reference itself cannot travel to the GPU box) as eager torch-CUDA
    ops on the full training shape.  A reported baseline next to cpu_baseline, not a product path."""
    import torch
    p, e, ga, gb = [t.to(dev) for t in cpu_inputs(N_ROWS)]
    cpu_port_step(p, e, ga, gb)
    torch.cuda.synchronize(dev)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    cpu_port_step(p, e, ga, gb)
    t1.record()
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1)
    return {"value": N_ROWS * T * S / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "note": "the reference's own functions (or the oracle port when absent) as eager torch-CUDA ops, "
                    "fp32, full training shape, 1 rep after 1 warm-up"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path -- the unmodified
    codebase/utils/monotonic_attention.py functions, shipped to the GPU box under baseline/_ref by
    oracle/ship_reference.py (`kind: "reference"`; the oracle port only if they are absent) -- on
    all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    best, mean, sec, cores = time_cpu(max(1, args.steps), max(1, min(args.warmup, 1)))
    kind = _reference_functions()[0]
    sample = (f"{CPU_SAMPLE_ROWS} of {N_ROWS} rows (full tgt {T} x src {S}), fp32, fwd+bwd through the "
              f"{'unmodified reference functions (baseline/_ref)' if kind == 'reference' else 'oracle port'}, "
              f"mean of {max(1, args.steps)}")
    line = {
        "impl": "reference", "metric": METRIC, "value": mean, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config(args.gpus),
        "cpu_baseline": {"value": mean, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mean, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import simulst_b200
    from simulst_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # The 6.3 MB gradient all-reduce overlaps the next step's kernels, which fill every SM
        # with 4 resident rows; each NCCL channel takes an SM's worth of resources away from
        # them.  4 channels keep the all-reduce hidden (1 channel exposes it: 0.76 ms/step at
        # N=4) and cost the kernels less than the default (0.530 vs 0.543 ms/step at N=4).
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    g = torch.Generator().manual_seed(1234 + rank)
    dt = torch.bfloat16
    # pinned host copies (the e2e leg uploads them every step)
    p_host = torch.sigmoid(torch.randn(N_ROWS, T, S, generator=g) - 2.0).to(dt).pin_memory()
    e_host = torch.randn(N_ROWS, T, S, generator=g).to(dt).pin_memory()
    p = p_host.to(dev, non_blocking=True)
    e = e_host.to(dev, non_blocking=True)
    alpha = torch.empty(N_ROWS, T, S, device=dev)
    beta = torch.empty_like(alpha)
    side = torch.empty(N_ROWS, T, 2, device=dev)
    ga = ((torch.arange(1, S + 1, device=dev).float() / S).expand(N_ROWS, T, S)
          + 1e-2 * torch.randn(N_ROWS, T, S, device=dev)).contiguous()
    gb = torch.randn(N_ROWS, T, S, device=dev)
    gp = torch.empty_like(p)
    ge = torch.empty_like(e)
    gp_host = torch.empty(N_ROWS, T, S, dtype=dt).pin_memory()
    ge_host = torch.empty(N_ROWS, T, S, dtype=dt).pin_memory()
    status = _lib.status_word(dev)
    grad_buf = torch.randn(GRAD_BUF_ELEMS, device=dev)
    # Gradient all-reduce: NVSwitch multicast kernel (simulst_multimem_allreduce_f32, a few CTAs) on a
    # symmetric buffer when the box supports it, NCCL otherwise (SIMULST_ALLREDUCE=nccl forces NCCL).
    mm = None
    mm_note = None
    if world > 1 and os.environ.get("SIMULST_ALLREDUCE", "multimem") == "multimem":
        try:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(GRAD_BUF_ELEMS, dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            if not hdl.multicast_ptr:
                raise RuntimeError("no multicast mapping")
            mm = {"hdl": hdl, "ptr": int(hdl.multicast_ptr), "stream": torch.cuda.Stream(),
                  "ev": torch.cuda.Event(), "ctas": int(os.environ.get("SIMULST_ALLREDUCE_CTAS", "2"))}
            grad_buf = buf
            # known-answer check of the kernel: every rank holds rank+1 -> the sum everywhere
            grad_buf.fill_(float(rank + 1))
            torch.cuda.synchronize()
            hdl.barrier(channel=0)
            _lib.check(lib.simulst_multimem_allreduce_f32(mm["ptr"], GRAD_BUF_ELEMS, rank, world, mm["ctas"],
                                                          torch.cuda.current_stream().cuda_stream),
                       "simulst_multimem_allreduce_f32")
            hdl.barrier(channel=1)
            torch.cuda.synchronize()
            want = world * (world + 1) / 2.0
            if not bool((grad_buf == want).all()):
                raise RuntimeError(f"multimem all-reduce check failed: {float(grad_buf.min())}..{float(grad_buf.max())} != {want}")
            grad_buf.normal_()
        except Exception as exc:
            mm = None
            mm_note = repr(exc)
            grad_buf = torch.randn(GRAD_BUF_ELEMS, device=dev)
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    flags = _lib.MMA_MASS_PRESERVATION | _lib.MMA_SOFT

    def fwd():
        rc = lib.simulst_mma_train_fwd(p.data_ptr(), _lib.BF16, e.data_ptr(), _lib.BF16, None,
                                       alpha.data_ptr(), beta.data_ptr(), side.data_ptr(),
                                       N_ROWS, T, S, EPS, 0, flags, status.data_ptr(), st)
        _lib.check(rc, "simulst_mma_train_fwd")

    def bwd():
        rc = lib.simulst_mma_train_bwd(p.data_ptr(), _lib.BF16, e.data_ptr(), _lib.BF16, None,
                                       alpha.data_ptr(), side.data_ptr(), ga.data_ptr(), gb.data_ptr(),
                                       gp.data_ptr(), _lib.BF16, ge.data_ptr(), _lib.BF16,
                                       N_ROWS, T, S, EPS, 0, flags, st)
        _lib.check(rc, "simulst_mma_train_bwd")

    pending = [None]

    def step(timers=None):
        if timers is not None:
            timers[0].record(stream)
        fwd()
        if timers is not None:
            timers[1].record(stream)
        bwd()
        if timers is not None:
            timers[2].record(stream)
        if world > 1:
            all_reduce_async()

    def all_reduce_async():
        """Sum the gradient buffer over the ranks, overlapped with the following kernels."""
        if mm is not None:
            mm["ev"].record(stream)
            mm["stream"].wait_event(mm["ev"])
            with torch.cuda.stream(mm["stream"]):
                mm["hdl"].barrier(channel=0)
                _lib.check(lib.simulst_multimem_allreduce_f32(mm["ptr"], GRAD_BUF_ELEMS, rank, world, mm["ctas"],
                                                              mm["stream"].cuda_stream),
                           "simulst_multimem_allreduce_f32")
                mm["hdl"].barrier(channel=1)
            pending[0] = mm
            return
        if pending[0] is not None:
            pending[0].wait()
        pending[0] = dist.all_reduce(grad_buf, async_op=True)

    def all_reduce_join():
        if pending[0] is None:
            return
        if mm is not None:
            stream.wait_stream(mm["stream"])
        else:
            pending[0].wait()
        pending[0] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing: W warm-up, exactly K timed steps
    for _ in range(args.warmup):
        step()
    barrier()
    simulst_b200.reset_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    for k in range(args.steps):
        step(ev[k])
    all_reduce_join()
    t_end.record(stream)
    barrier()
    launches = simulst_b200.launch_count()
    ms_total = t_begin.elapsed_time(t_end)
    t_max = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_step = float(t_max.item()) / args.steps
    fwd_ms = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)) / args.steps
    bwd_ms = sum(ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)) / args.steps
    simulst_b200.check_status(dev)
    elems = N_ROWS * T * S
    value = elems * world / (ms_step * 1e-3)

    # ---------------- end to end through the public host-buffer API (copies inside the region):
    # simulst_b200.host_pipeline.MMAHostPipeline streams row chunks H2D -> fwd+bwd -> D2H
    from simulst_b200.host_pipeline import MMAHostPipeline
    pipe = MMAHostPipeline(N_ROWS, T, S, dtype=dt, device=dev, chunks=args.e2e_chunks,
                           compute_streams=args.e2e_streams, eps=EPS, mass_preservation=True, soft=True)

    def e2e_step():
        pipe.step(p_host, e_host, ga, gb, gp_host, ge_host)
        if world > 1:
            all_reduce_async()

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    simulst_b200.reset_launch_count()
    t_begin.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    all_reduce_join()
    t_end.record(stream)
    barrier()
    e2e_launches = simulst_b200.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    e_max = torch.tensor([t_begin.elapsed_time(t_end)], device=dev)
    if world > 1:
        dist.all_reduce(e_max, op=dist.ReduceOp.MAX)
    e2e_ms = float(e_max.item()) / e2e_steps
    h2d = pipe.h2d_bytes_per_step
    d2h = pipe.d2h_bytes_per_step
    simulst_b200.check_status(dev)

    extras = {}
    cpu_base = None
    if rank == 0 and world == 1:
        # expected-delay epilogue on the same resident buffers: the latency
        # loss reaches alpha only through the [N,T] delays, so the backward reads no grad_alpha
        try:
            delays = torch.empty(N_ROWS, T, device=dev)
            g_del = torch.randn(N_ROWS, T, device=dev) / S

            def step_delays():
                rc = lib.simulst_mma_train_fwd_delays(p.data_ptr(), _lib.BF16, e.data_ptr(), _lib.BF16, None,
                                                      alpha.data_ptr(), beta.data_ptr(), side.data_ptr(),
                                                      delays.data_ptr(), N_ROWS, T, S, EPS, 0, flags,
                                                      status.data_ptr(), st)
                _lib.check(rc, "simulst_mma_train_fwd_delays")
                rc = lib.simulst_mma_train_bwd_delays(p.data_ptr(), _lib.BF16, e.data_ptr(), _lib.BF16, None,
                                                      alpha.data_ptr(), side.data_ptr(), None, gb.data_ptr(),
                                                      g_del.data_ptr(), gp.data_ptr(), _lib.BF16, ge.data_ptr(),
                                                      _lib.BF16, N_ROWS, T, S, EPS, 0, flags, st)
                _lib.check(rc, "simulst_mma_train_bwd_delays")

            for _ in range(3):
                step_delays()
            torch.cuda.synchronize()
            d0 = torch.cuda.Event(enable_timing=True)
            d1 = torch.cuda.Event(enable_timing=True)
            d0.record(stream)
            for _ in range(10):
                step_delays()
            d1.record(stream)
            torch.cuda.synchronize()
            d_ms = d0.elapsed_time(d1) / 10
            peak_d, _ = measured_peak()
            extras["latency_epilogue"] = {
                "metric": "mma_fwd_bwd_with_expected_delays_elements_per_s", "value": elems / (d_ms * 1e-3),
                "unit": UNIT, "ms_per_step": d_ms,
                "algorithmic_bytes_per_element": BYTES_FWD + BYTES_BWD - 4,
                "roofline_frac": elems * (BYTES_FWD + BYTES_BWD - 4) / (d_ms * 1e-3) / 1e9 / peak_d,
                "note": "simulst_mma_train_fwd_delays + simulst_mma_train_bwd_delays(grad_alpha=NULL): "
                        "expected delays as a by-product of the forward scan, their gradient injected in "
                        "the backward kernel (mma_criterion.py:146-157)"}
        except Exception as exc:  # pragma: no cover
            extras["latency_epilogue"] = {"error": repr(exc)}
        # right-padded batch (SURVEY 8d variant ii: lengths ~ U[S/2, S]) with and without the
        # caller's right-padding promise (simulst_b200.assume_right_padding / SIMULST_MMA_RIGHT_PADDING)
        try:
            gm = torch.Generator().manual_seed(1236)
            lens = torch.randint(S // 2, S + 1, (N_ROWS,), generator=gm)
            mask = (torch.arange(S)[None, :] >= lens[:, None]).to(dev).view(torch.uint8).contiguous()
            res = {}
            for name, fl, split in (("single_pass_arbitrary_mask_kernels", flags, 0),
                                    ("default_two_passes_split_by_row", flags, 1),
                                    ("right_padding_promise", flags | _lib.MMA_RIGHT_PADDING, 1)):
                lib.simulst_mma_set_mask_split(split)

                def step_masked():
                    rc = lib.simulst_mma_train_fwd(p.data_ptr(), _lib.BF16, e.data_ptr(), _lib.BF16, mask.data_ptr(),
                                                   alpha.data_ptr(), beta.data_ptr(), side.data_ptr(),
                                                   N_ROWS, T, S, EPS, 0, fl, status.data_ptr(), st)
                    _lib.check(rc, "simulst_mma_train_fwd")
                    rc = lib.simulst_mma_train_bwd(p.data_ptr(), _lib.BF16, e.data_ptr(), _lib.BF16, mask.data_ptr(),
                                                   alpha.data_ptr(), side.data_ptr(), ga.data_ptr(), gb.data_ptr(),
                                                   gp.data_ptr(), _lib.BF16, ge.data_ptr(), _lib.BF16,
                                                   N_ROWS, T, S, EPS, 0, fl, st)
                    _lib.check(rc, "simulst_mma_train_bwd")
                for _ in range(3):
                    step_masked()
                torch.cuda.synchronize()
                m0 = torch.cuda.Event(enable_timing=True)
                m1 = torch.cuda.Event(enable_timing=True)
                m0.record(stream)
                for _ in range(10):
                    step_masked()
                m1.record(stream)
                torch.cuda.synchronize()
                res[name] = {"ms_per_step": m0.elapsed_time(m1) / 10,
                             "value": elems / (m0.elapsed_time(m1) / 10 * 1e-3), "unit": UNIT}
            lib.simulst_mma_set_mask_split(1)
            extras["masked_batch"] = {"config": "training shape, right-padded source lengths ~ U[S/2, S] (seed 1236), "
                                                "elements counted over the full [N,T,S] grid", **res}
        except Exception as exc:  # pragma: no cover
            extras["masked_batch"] = {"error": repr(exc)}
        try:
            extras["long_rows_small_batch"] = bench_cluster(lib, dev)
        except Exception as exc:  # pragma: no cover
            extras["long_rows_small_batch"] = {"error": repr(exc)}
        try:
            extras["unaligned_rows"] = bench_unaligned(lib, dev)
        except Exception as exc:  # pragma: no cover
            extras["unaligned_rows"] = {"error": repr(exc)}
        extras.update(side_benchmarks(lib, dev))
        best, mean, sec, cores = time_cpu(reps=2, warmup=1)
        cpu_base = {"value": mean, "unit": UNIT, "cores": cores, "kind": _reference_functions()[0],
                    "sample": f"{CPU_SAMPLE_ROWS} of {N_ROWS} rows (full tgt {T} x src {S}), fp32, "
                              f"fwd+bwd, mean of 2 after 1 warm-up ({sec:.2f} s each)"}
        # the tensors the timed region produced, checked against the reference on two rows
        fwd()
        bwd()
        torch.cuda.synchronize()
        rows = [0, N_ROWS - 1]
        cpu_base["parity_check"] = {
            "rows": rows, "against": cpu_base["kind"],
            **parity_rows(p_host[rows], e_host[rows], ga[rows].cpu(), gb[rows].cpu(),
                          (alpha[rows], beta[rows], gp[rows], ge[rows]))}
        try:
            del gp_host, ge_host
            torch.cuda.empty_cache()
            cpu_base["reference_gpu_eager"] = time_gpu_eager(dev)
        except Exception as exc:  # pragma: no cover
            cpu_base["reference_gpu_eager"] = {"error": repr(exc)}

    if rank == 0:
        peak, peak_src = measured_peak()
        ach = elems * BYTES_BWD / (bwd_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config(world),
            "roofline": {"bound": "hbm", "kernel": "mma_bwd_fast_kernel", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic("mma_bwd_fast_kernel"),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": elems * BYTES_BWD,
                         "kernel_ms": bwd_ms,
                         "fwd": {"kernel": "mma_fwd_pipe_kernel", "kernel_ms": fwd_ms,
                                 "achieved": elems * BYTES_FWD / (fwd_ms * 1e-3) / 1e9,
                                 "frac": elems * BYTES_FWD / (fwd_ms * 1e-3) / 1e9 / peak,
                                 "traffic": ncu_traffic("mma_fwd_pipe_kernel")},
                         "fwd_bwd_frac": elems * (BYTES_FWD + BYTES_BWD) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / peak},
            "e2e": {"value": elems * world / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "steps": e2e_steps, "api": "simulst_b200.host_pipeline.MMAHostPipeline.step",
                    "row_chunks": len(pipe.bounds), "compute_streams": len(pipe.s_comp),
                    "gpu_launches": int(e2e_launches),
                    "host_link_gbs_all_ranks": (h2d + d2h) * world / (e2e_ms * 1e-3) / 1e9,
                    "note": "PCIe / host-memory bound: 268 MB up + 268 MB down per rank and step; one rank alone "
                            "moves ~87 GB/s (both directions together), all ranks of one box share the host's "
                            "root complexes and memory (DESIGN.md 6)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world > 1:
            line["all_reduce"] = ({"impl": "simulst_multimem_allreduce_f32 (NVSwitch multicast ld_reduce + st), "
                                           f"{mm['ctas']} CTAs, symmetric buffer {GRAD_BUF_ELEMS * 4 / 1e6:.1f} MB, "
                                           "overlapped with the next step"}
                                  if mm is not None else
                                  {"impl": "NCCL all_reduce (async), NCCL_MAX_NCHANNELS=" +
                                           os.environ.get("NCCL_MAX_NCHANNELS", "default"),
                                   "multimem_unavailable": mm_note})
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def side_benchmarks(lib, dev):
    """CIF (BASELINE config 3) and the incremental step (config 4): reported next to the headline,
    N=1 only, not part of `value`."""
    import torch
    from simulst_b200 import ops
    from simulst_b200.models.torch_cif import cif_function
    out = {}
    try:
        g = torch.Generator().manual_seed(2024)
        b, s, c = 64, 1500, 256
        x = torch.randn(b, s, c, generator=g).to(dev).requires_grad_()
        a = torch.sigmoid(torch.randn(b, s, generator=g) - 1.0).to(dev).requires_grad_()
        tl = a.detach().sum(1).round().clamp(min=1).long()
        res = cif_function(x, a, beta=1.0, tail_thres=0.5, target_lengths=tl)
        go = torch.randn_like(res["cif_out"][0])
        gd = torch.randn_like(res["delays"][0])

        def cif_step():
            x.grad = None
            a.grad = None
            r = cif_function(x, a, beta=1.0, tail_thres=0.5, target_lengths=tl)
            torch.autograd.backward([r["cif_out"][0], r["delays"][0]], [go, gd])

        for _ in range(3):
            cif_step()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        reps = 20
        t0.record()
        for _ in range(reps):
            cif_step()
        t1.record()
        torch.cuda.synchronize()
        api_ms = t0.elapsed_time(t1) / reps
        # the same call with the target lengths still on the HOST (an extension: T is known without the
        # device read the reference's return type otherwise forces, so nothing synchronises)
        tl_host = tl.cpu()

        def cif_step_host():
            x.grad = None
            a.grad = None
            r = cif_function(x, a, beta=1.0, tail_thres=0.5, target_lengths=tl_host)
            torch.autograd.backward([r["cif_out"][0], r["delays"][0]], [go, gd])
        for _ in range(3):
            cif_step_host()
        torch.cuda.synchronize()
        t0.record()
        for _ in range(reps):
            cif_step_host()
        t1.record()
        torch.cuda.synchronize()
        api_host_ms = t0.elapsed_time(t1) / reps
        t_out = int(res["cif_out"][0].shape[1])
        alg = b * s * (c * 4 * (3 + 2 * t_out / s) + 8)
        peak, _ = measured_peak()

        # the same step through the C ABI on caller-allocated, HBM-resident buffers
        # (plan + forward + backward = 4 launches, no host read: T is known from target_lengths)
        from simulst_b200 import _lib
        xd, ad = x.detach(), a.detach()
        desired = (1.0 * tl.float() + 1e-4).contiguous()
        csum = torch.empty(b, s, device=dev)
        scale = torch.empty(b, device=dev)
        asum = torch.empty(b, device=dev)
        len0 = torch.empty(b, dtype=torch.int64, device=dev)
        cnt = torch.zeros(2, dtype=torch.int32, device=dev)
        seg = torch.empty(b, t_out + 2, dtype=torch.int32, device=dev)
        o = torch.empty(b, t_out, c, device=dev)
        dl = torch.empty(b, t_out, device=dev)
        gx = torch.empty_like(xd)
        gal = torch.empty_like(ad)
        ws = torch.empty(2 * b * s, device=dev)
        status = _lib.status_word(dev)
        st = torch.cuda.current_stream().cuda_stream

        def cif_abi():
            rc = lib.simulst_cif_plan(ad.data_ptr(), 0, None, desired.data_ptr(), tl.data_ptr(),
                                      csum.data_ptr(), scale.data_ptr(), asum.data_ptr(), len0.data_ptr(),
                                      cnt.data_ptr(), seg.data_ptr(), t_out + 2, b, s, 1.0,
                                      status.data_ptr(), st)
            _lib.check(rc, "simulst_cif_plan")
            rc = lib.simulst_cif_fwd(xd.data_ptr(), 0, csum.data_ptr(), scale.data_ptr(), ad.data_ptr(), 0,
                                     None, seg.data_ptr(), t_out + 2, o.data_ptr(), dl.data_ptr(), None,
                                     len0.data_ptr(), None, None, b, s, c, t_out, t_out, 1.0, 0.5, 1, st)
            _lib.check(rc, "simulst_cif_fwd")
            rc = lib.simulst_cif_bwd(xd.data_ptr(), 0, csum.data_ptr(), scale.data_ptr(), ad.data_ptr(), 0,
                                     None, go.data_ptr(), gd.data_ptr(), None, None, None, asum.data_ptr(),
                                     None, gx.data_ptr(), gal.data_ptr(), ws.data_ptr(), b, s, c, t_out,
                                     t_out, 1.0, 0.5, 1, st)
            _lib.check(rc, "simulst_cif_bwd")

        for _ in range(3):
            cif_abi()
        torch.cuda.synchronize()
        reps = 50
        t0.record()
        for _ in range(reps):
            cif_abi()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        assert torch.equal(o, res["cif_out"][0].detach()), "C-ABI and Python API CIF outputs differ"
        # cold: L2 (126 MB) flushed before every step, so the three kernels' inputs come from HBM
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        cold = []
        for _ in range(10):
            flush.zero_()
            t0.record()
            cif_abi()
            t1.record()
            torch.cuda.synchronize()
            cold.append(t0.elapsed_time(t1))
        del flush
        cold_ms = sorted(cold)[len(cold) // 2]
        # same-run CPU baseline: the reference's cif_function + autograd on this box's cores
        # (harness shape of codebase/models/torch_cif/benchmark.py:87-128)
        cif_cpu = None
        try:
            from oracle import ref_loader
            if ref_loader.available():
                ref_cif, kind = ref_loader.load_cif().cif_function, "reference"
            else:
                from oracle import cif as ocif
                ref_cif, kind = ocif.cif_function, "port"
            torch.set_num_threads(os.cpu_count() or 1)
            xc, ac = x.detach().cpu().requires_grad_(), a.detach().cpu().requires_grad_()
            tlc, goc, gdc = tl.cpu(), go.cpu(), gd.cpu()
            best = float("inf")
            for k in range(3):
                xc.grad = None
                ac.grad = None
                c0 = time.perf_counter()
                r = ref_cif(xc, ac, beta=1.0, tail_thres=0.5, target_lengths=tlc)
                torch.autograd.backward([r["cif_out"][0], r["delays"][0]], [goc, gdc])
                if k > 0:
                    best = min(best, time.perf_counter() - c0)
            cif_cpu = {"value": b * s / best, "unit": "frames/s", "ms_per_step": best * 1e3,
                       "cores": torch.get_num_threads(), "kind": kind,
                       "sample": "full config (B=64 S=1500 C=256 fp32 fwd+bwd), best of 2 after 1 warm-up"}
        except Exception as exc:  # pragma: no cover
            cif_cpu = {"error": repr(exc)}
        out["cif"] = {"metric": "cif_fwd_bwd_frames_per_s", "value": b * s / (ms * 1e-3), "unit": "frames/s",
                      "config": f"B={b} S={s} C={c} fp32 beta=1.0 training mode, T={t_out}; working set "
                                f"{alg / 1e6:.0f} MB > L2",
                      "ms_per_step": ms, "algorithmic_bytes": alg,
                      "roofline_frac": alg / (cold_ms * 1e-3) / 1e9 / peak,
                      "ms_per_step_cold_l2": cold_ms,
                      "roofline_frac_warm_l2": alg / (ms * 1e-3) / 1e9 / peak,
                      "timing": "roofline_frac: L2 flushed before every step (median of 10); ms_per_step / "
                                "value / roofline_frac_warm_l2: 50 back-to-back steps (the chain re-hits L2)",
                      "cpu_baseline": cif_cpu,
                      "path": "C ABI, buffers resident: simulst_cif_plan + simulst_cif_fwd + simulst_cif_bwd",
                      "traffic": sum(ncu_traffic(k) or 0 for k in ("cif_plan_kernel", "cif_fwd_tile_kernel",
                                                                    "cif_bwd_tile_kernel", "cif_bwd_alpha_kernel")) or None,
                      "python_api": {"value": b * s / (api_ms * 1e-3), "ms_per_step": api_ms,
                                     "roofline_frac": alg / (api_ms * 1e-3) / 1e9 / peak,
                                     "note": "cif_function + autograd backward incl. allocations and the "
                                             "reference's host read of T (cif.py:72)",
                                     "host_resident_target_lengths": {
                                         "value": b * s / (api_host_ms * 1e-3), "ms_per_step": api_host_ms,
                                         "roofline_frac": alg / (api_host_ms * 1e-3) / 1e9 / peak,
                                         "note": "target_lengths passed as a CPU tensor: no device read, no sync"}}}
    except Exception as exc:  # pragma: no cover
        out["cif"] = {"error": repr(exc)}
    try:
        g = torch.Generator().manual_seed(3000)
        r, s = 256 * 4, 1024
        p = torch.sigmoid(torch.randn(r, s, generator=g) - 2.0).to(dev)
        se = torch.randn(r, s, generator=g).to(dev)
        hs = torch.zeros(r, dtype=torch.long, device=dev)
        for _ in range(3):
            ops.mma_step(p, hs, se, None, True)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        reps = 50
        hs.zero_()
        t0.record()
        for _ in range(reps):
            ops.mma_step(p, hs, se, None, True)
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) / reps * 1e3
        out["incremental_step"] = {"metric": "mma_infer_step_us_per_layer", "value": us, "unit": "us",
                                   "config": f"256 utterances x 4 heads, src {s}, infinite lookback, fp32",
                                   "rows_per_s": r / (us * 1e-6)}
        out["incremental_step"].update(step_variants(lib, dev, p, se))
        out["incremental_step"]["cpu_baseline"] = step_cpu_baseline(p.cpu(), se.cpu())
    except Exception as exc:  # pragma: no cover
        out["incremental_step"] = {"error": repr(exc)}
    for name, fn in (("config1_forward", bench_config1), ("pooled_p_choose", bench_pooled),
                     ("ssnt_loss", bench_ssnt), ("ctc_best_alignment", bench_ctc), ("latency_dal", bench_dal)):
        try:
            out[name] = fn(lib, dev)
        except Exception as exc:  # pragma: no cover
            out[name] = {"error": repr(exc)}
    return out


def _events_us(fn, reps, flush=None):
    """Median microseconds of `fn` over `reps` CUDA-event timings (L2 flushed before each when a
    flush buffer is given)."""
    import torch
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        fn()
        t1.record()
        torch.cuda.synchronize()
        ts.append(t0.elapsed_time(t1) * 1e3)
    return sorted(ts)[len(ts) // 2]


def step_variants(lib, dev, p, se):
    """The decoding step as the agent runs it: 6 decoder layers per target token
    (models/mma_model.py:191-210 calls the attention once per layer).  (a) six C-ABI launches on
    preallocated buffers, (b) the same six launches replayed from one CUDA graph."""
    import torch
    from simulst_b200 import _lib
    layers = 6
    r, s = p.shape
    hs = [torch.zeros(r, dtype=torch.long, device=dev) for _ in range(layers)]
    hr = [torch.empty(r, dtype=torch.uint8, device=dev) for _ in range(layers)]
    al = [torch.empty(r, s, device=dev) for _ in range(layers)]
    be = [torch.empty(r, s, device=dev) for _ in range(layers)]

    def six(stream_ptr):
        for k in range(layers):
            rc = lib.simulst_mma_step(p.data_ptr(), _lib.F32, se.data_ptr(), _lib.F32, None, hs[k].data_ptr(),
                                      hr[k].data_ptr(), al[k].data_ptr(), be[k].data_ptr(), r, s,
                                      _lib.MMA_MASS_PRESERVATION | _lib.MMA_SOFT, stream_ptr)
            _lib.check(rc, "simulst_mma_step")

    cur = torch.cuda.current_stream().cuda_stream
    six(cur)
    torch.cuda.synchronize()
    abi_us = _events_us(lambda: six(cur), 30)
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        six(side.cuda_stream)
    side.synchronize()
    with torch.cuda.graph(graph):
        six(torch.cuda.current_stream().cuda_stream)
    graph.replay()
    torch.cuda.synchronize()
    graph_us = _events_us(graph.replay, 30)
    return {"six_layers_c_abi_us": abi_us, "six_layers_cuda_graph_us": graph_us,
            "per_layer_c_abi_us": abi_us / layers, "per_layer_cuda_graph_us": graph_us / layers}


def step_cpu_baseline(p, se):
    """The reference's monotonic_attention_process_infer body
    (modules/monotonic_multihead_attention.py:152-299) on this box's cores: the real class with
    its two projection calls (p_choose, energy_from_qk) returning the precomputed tensors, so the
    timed work is the same as the kernel's."""
    import torch
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return {"unavailable": "reference files not present"}
        heads = 4
        r, s = p.shape
        bsz = r // heads
        att = ref_loader.make_attention("infinite_lookback", 8 * heads, heads).eval()
        att.p_choose = lambda q, k, m=None, inc=None: p.unsqueeze(1)
        att.energy_from_qk = lambda q, k, kind, key_padding_mask=None, bias=0: se.unsqueeze(1)
        q = torch.zeros(1, bsz, 8 * heads)
        k = torch.zeros(s, bsz, 8 * heads)
        torch.set_num_threads(os.cpu_count() or 1)
        inc = {}
        best = float("inf")
        with torch.no_grad():
            for it in range(6):
                c0 = time.perf_counter()
                att.monotonic_attention_process_infer(q, k, None, inc)
                if it > 0:
                    best = min(best, time.perf_counter() - c0)
        return {"value": best * 1e6, "unit": "us", "cores": torch.get_num_threads(), "kind": "reference",
                "sample": "same 1024 rows x src 1024, 5 consecutive steps after 1 warm-up, best"}
    except Exception as exc:  # pragma: no cover
        return {"error": repr(exc)}


def bench_cluster(lib, dev):
    """SURVEY's long-form row count (BASELINE config 5: 8 utterances x 8 heads = 64 rows), forward, tgt 128: one CTA
    per row (84 SMs idle) against one thread-block cluster per row (csrc/mma_fwd_cluster.cuh: the library's choice
    for <= 74 rows of more than 2560 frames)."""
    import torch
    from simulst_b200 import _lib
    st = torch.cuda.current_stream(dev).cuda_stream
    n, t = 64, 128
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    peak, _ = measured_peak()
    out = {"config": f"{n} rows x tgt {t}, bf16 in, forward (alpha + beta), L2 flushed before every launch, median of 7",
           "algorithmic_bytes_per_element": 12}
    g = torch.Generator().manual_seed(5)
    for s_len in (3000, 4096, 6000, 8192):
        p = torch.sigmoid(torch.randn(n, t, s_len, generator=g) - 2).to(dev, torch.bfloat16)
        e = torch.randn(n, t, s_len, generator=g).to(dev, torch.bfloat16)
        alpha = torch.empty(n, t, s_len, device=dev)
        beta = torch.empty_like(alpha)
        side = torch.empty(n, t, 2, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        fl = _lib.MMA_SOFT | _lib.MMA_MASS_PRESERVATION

        def fwd():
            _lib.check(lib.simulst_mma_train_fwd(p.data_ptr(), _lib.BF16, e.data_ptr(), _lib.BF16, None, alpha.data_ptr(),
                                                 beta.data_ptr(), side.data_ptr(), n, t, s_len, EPS, 0, fl,
                                                 status.data_ptr(), st), "simulst_mma_train_fwd")
        res = {}
        try:
            for mode, name in ((0, "one_cta_per_row_us"), (1, "cluster_per_row_us")):
                lib.simulst_mma_set_cluster(mode)
                fwd()
                torch.cuda.synchronize()
                res[name] = _events_us(fwd, 7, flush)
        finally:
            lib.simulst_mma_set_cluster(1)
        res["speedup"] = res["one_cta_per_row_us"] / res["cluster_per_row_us"]
        res["roofline_frac"] = n * t * s_len * 12 / (res["cluster_per_row_us"] * 1e-6) / 1e9 / peak
        out[f"src{s_len}"] = res
        del p, e, alpha, beta
        torch.cuda.empty_cache()
    return out


def bench_unaligned(lib, dev):
    """Source lengths whose rows are not 16-byte multiples (S = 1500 in bf16 is the CIF config's own S; S = 999
    has rows at 2-byte offsets), 512 rows x 128 steps, fwd + bwd through the row-pitch entry points: outputs with
    the pitch simulst_mma_out_pitch(S) asks for (what the Python wrapper allocates; dense kernels with shifted
    staging), the same call with dense outputs (generic kernels), and the aligned neighbour S = 1504 / 1000."""
    import torch
    from simulst_b200 import _lib
    st = torch.cuda.current_stream(dev).cuda_stream
    n, t = N_ROWS, T
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {"config": f"{n} rows x tgt {t}, bf16 in, L2 flushed before every launch, median of 5",
           "algorithmic_bytes_per_element": 32}
    peak, _ = measured_peak()
    g = torch.Generator().manual_seed(77)
    for s_len, pitched in ((1500, True), (1500, False), (1504, True), (999, True), (999, False), (1000, True)):
        ld = int(lib.simulst_mma_out_pitch(s_len)) if pitched else s_len
        p = torch.sigmoid(torch.randn(n, t, s_len, generator=g) - 2).to(dev, torch.bfloat16)
        e = torch.randn(n, t, s_len, generator=g).to(dev, torch.bfloat16)
        alpha = torch.empty(n, t, ld, device=dev)
        beta = torch.empty_like(alpha)
        side = torch.empty(n, t, 2, device=dev)
        ga = torch.randn(n, t, s_len, device=dev) * 0.01
        gb = torch.randn(n, t, s_len, device=dev)
        gp = torch.empty(n, t, ld, device=dev, dtype=torch.bfloat16)
        ge = torch.empty_like(gp)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        fl = _lib.MMA_SOFT | _lib.MMA_MASS_PRESERVATION

        def fwd():
            _lib.check(lib.simulst_mma_train_fwd_pitched(
                p.data_ptr(), _lib.BF16, s_len, e.data_ptr(), _lib.BF16, s_len, None, alpha.data_ptr(), ld,
                beta.data_ptr(), ld, side.data_ptr(), None, n, t, s_len, EPS, 0, fl, status.data_ptr(), st),
                "simulst_mma_train_fwd_pitched")

        def bwd():
            _lib.check(lib.simulst_mma_train_bwd_pitched(
                p.data_ptr(), _lib.BF16, s_len, e.data_ptr(), _lib.BF16, s_len, None, alpha.data_ptr(), ld,
                side.data_ptr(), ga.data_ptr(), s_len, gb.data_ptr(), s_len, None, gp.data_ptr(), _lib.BF16, ld,
                ge.data_ptr(), _lib.BF16, ld, n, t, s_len, EPS, 0, fl, st), "simulst_mma_train_bwd_pitched")
        fwd()
        bwd()
        torch.cuda.synchronize()
        f_us, b_us = _events_us(fwd, 5, flush), _events_us(bwd, 5, flush)
        el = n * t * s_len
        out[f"src{s_len}_" + ("pitch%d" % ld if pitched else "dense_outputs")] = {
            "fwd_us": f_us, "bwd_us": b_us, "value": el / ((f_us + b_us) * 1e-6), "unit": UNIT,
            "roofline_frac": el * 32 / ((f_us + b_us) * 1e-6) / 1e9 / peak}
        del p, e, alpha, beta, ga, gb, gp, ge
        torch.cuda.empty_cache()
    return out


def bench_config1(lib, dev):
    """BASELINE config 1: infinite-lookback expected alignment + soft attention FORWARD, fp32,
    B=8 H=4 tgt=32 src=256 -- the reference's own CPU-runnable case.  262 144 elements x 16 B =
    4.2 MB: launch/latency bound on the GPU (0.65 us at HBM peak), reported as it is."""
    import torch
    from simulst_b200 import _lib
    n, t, s = 32, 32, 256
    g = torch.Generator().manual_seed(1234)
    p_c = torch.sigmoid(torch.randn(n, t, s, generator=g) - 2.0)
    e_c = torch.randn(n, t, s, generator=g)
    p, e = p_c.to(dev), e_c.to(dev)
    alpha, beta = torch.empty(n, t, s, device=dev), torch.empty(n, t, s, device=dev)
    status = _lib.status_word(dev)
    st = torch.cuda.current_stream().cuda_stream
    flags = _lib.MMA_MASS_PRESERVATION | _lib.MMA_SOFT

    def fwd():
        rc = lib.simulst_mma_train_fwd(p.data_ptr(), _lib.F32, e.data_ptr(), _lib.F32, None, alpha.data_ptr(),
                                       beta.data_ptr(), None, n, t, s, EPS, 0, flags, status.data_ptr(), st)
        _lib.check(rc, "simulst_mma_train_fwd")
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    us = _events_us(fwd, 20, flush)
    del flush
    kind, align, mass, soft = _reference_functions()
    torch.set_num_threads(os.cpu_count() or 1)
    best = float("inf")
    with torch.no_grad():
        for k in range(4):
            c0 = time.perf_counter()
            a_r = mass(align(p_c.float(), None, eps=EPS), None)
            b_r = soft(a_r, e_c, padding_mask=None, chunk_size=None, eps=EPS)
            if k > 0:
                best = min(best, time.perf_counter() - c0)
    err_a = float((alpha.cpu() - a_r).abs().max())
    err_b = float((beta.cpu() - b_r).abs().max())
    if not (err_a <= 2e-6 and err_b <= 2e-6):
        raise SystemExit(f"config 1 parity check failed: max|d alpha| {err_a:.3e}, max|d beta| {err_b:.3e}")
    peak, _ = measured_peak()
    elems = n * t * s
    return {"metric": "mma_expected_alignment_fwd_elements_per_s", "value": elems / (us * 1e-6), "unit": UNIT,
            "config": "BASELINE config 1: B=8 H=4 tgt=32 src=256 fp32, forward, L2 flushed before each launch",
            "us_per_call": us, "algorithmic_bytes": elems * 16,
            "roofline_frac": elems * 16 / (us * 1e-6) / 1e9 / peak,
            "note": "4.2 MB per call: launch/latency bound (0.65 us at HBM peak)",
            "max_abs_err_vs_cpu": {"alpha": err_a, "beta": err_b},
            "cpu_baseline": {"value": elems / best, "unit": UNIT, "ms_per_call": best * 1e3,
                             "cores": torch.get_num_threads(), "kind": kind,
                             "sample": "full config, best of 3 after 1 warm-up"}}


def bench_pooled(lib, dev):
    """the training shape as exp/2-mma.sh:56-57 really runs it -- fixed pre-decision
    ratio 8, p_choose arrives POOLED [N,T,S/8] -- through simulst_mma_train_{fwd,bwd}_pooled (the
    pooled-grid kernels of csrc/mma_sparse.cu), next to the dense entry points fed the
    zero-upsampled tensor (what the reference computes).  Two modes: (a) the reference's full
    return values (dense alpha out, dense grad_alpha in); (b) latency-loss mode: alpha reaches the
    caller only through beta and the [N,T] expected delays (mma_criterion.py:146-157), so the dense
    alpha is neither written nor its gradient read."""
    import torch
    from simulst_b200 import _lib
    ratio = 8
    sp = S // ratio
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(4321)
    pp = torch.sigmoid(torch.randn(N_ROWS, T, sp, generator=g) - 1.0).to(dev, dt)
    e = torch.randn(N_ROWS, T, S, generator=g).to(dev, dt)
    pd = torch.zeros(N_ROWS, T, S, device=dev, dtype=dt)
    pd[:, :, ratio - 1::ratio] = pp
    alpha = torch.empty(N_ROWS, T, S, device=dev)
    beta = torch.empty_like(alpha)
    side = torch.empty(N_ROWS, T, 2, device=dev)
    delays = torch.empty(N_ROWS, T, device=dev)
    ga = torch.randn(N_ROWS, T, S, device=dev) * 1e-2
    gb = torch.randn(N_ROWS, T, S, device=dev)
    gd = torch.randn(N_ROWS, T, device=dev) / S
    gpp, ge, gpd = torch.empty_like(pp), torch.empty_like(e), torch.empty_like(pd)
    ws = torch.empty(int(lib.simulst_mma_pooled_workspace_bytes(N_ROWS, T, S, ratio)), dtype=torch.uint8, device=dev)
    status = _lib.status_word(dev)
    st = torch.cuda.current_stream().cuda_stream
    flags = _lib.MMA_MASS_PRESERVATION | _lib.MMA_SOFT
    B16 = _lib.BF16

    gm = torch.Generator().manual_seed(1236)
    lens = torch.randint(S // 2, S + 1, (N_ROWS,), generator=gm)
    mask = (torch.arange(S)[None, :] >= lens[:, None]).to(dev).view(torch.uint8).contiguous()

    def step_pooled(lean, masked=False):
        mk = mask.data_ptr() if masked else None
        fl = flags | (_lib.MMA_RIGHT_PADDING if masked else 0)
        if masked:
            rc = lib.simulst_mma_train_fwd_pooled(pp.data_ptr(), B16, ratio, e.data_ptr(), B16, mk, None,
                                                  alpha.data_ptr(), beta.data_ptr(), side.data_ptr(), None,
                                                  ws.data_ptr(), N_ROWS, T, S, EPS, 0, fl, status.data_ptr(), st)
            _lib.check(rc, "simulst_mma_train_fwd_pooled")
            rc = lib.simulst_mma_train_bwd_pooled(pp.data_ptr(), B16, ratio, e.data_ptr(), B16, mk, None,
                                                  None, side.data_ptr(), ga.data_ptr(), gb.data_ptr(), None,
                                                  gpp.data_ptr(), B16, None, ge.data_ptr(), B16, ws.data_ptr(),
                                                  N_ROWS, T, S, EPS, 0, fl, st)
            _lib.check(rc, "simulst_mma_train_bwd_pooled")
            return
        rc = lib.simulst_mma_train_fwd_pooled(pp.data_ptr(), B16, ratio, e.data_ptr(), B16, None, None,
                                              None if lean else alpha.data_ptr(), beta.data_ptr(), side.data_ptr(),
                                              delays.data_ptr() if lean else None, ws.data_ptr(),
                                              N_ROWS, T, S, EPS, 0, flags, status.data_ptr(), st)
        _lib.check(rc, "simulst_mma_train_fwd_pooled")
        rc = lib.simulst_mma_train_bwd_pooled(pp.data_ptr(), B16, ratio, e.data_ptr(), B16, None, None,
                                              None, side.data_ptr(), None if lean else ga.data_ptr(), gb.data_ptr(),
                                              gd.data_ptr() if lean else None, gpp.data_ptr(), B16, None,
                                              ge.data_ptr(), B16, ws.data_ptr(), N_ROWS, T, S, EPS, 0, flags, st)
        _lib.check(rc, "simulst_mma_train_bwd_pooled")

    def step_dense(lean, masked=False):
        if masked:
            fl = flags | _lib.MMA_RIGHT_PADDING
            rc = lib.simulst_mma_train_fwd(pd.data_ptr(), B16, e.data_ptr(), B16, mask.data_ptr(), alpha.data_ptr(),
                                           beta.data_ptr(), side.data_ptr(), N_ROWS, T, S, EPS, 0, fl,
                                           status.data_ptr(), st)
            _lib.check(rc, "simulst_mma_train_fwd")
            rc = lib.simulst_mma_train_bwd(pd.data_ptr(), B16, e.data_ptr(), B16, mask.data_ptr(), alpha.data_ptr(),
                                           side.data_ptr(), ga.data_ptr(), gb.data_ptr(), gpd.data_ptr(), B16,
                                           ge.data_ptr(), B16, N_ROWS, T, S, EPS, 0, fl, st)
            _lib.check(rc, "simulst_mma_train_bwd")
            return
        rc = lib.simulst_mma_train_fwd_delays(pd.data_ptr(), B16, e.data_ptr(), B16, None, alpha.data_ptr(),
                                              beta.data_ptr(), side.data_ptr(), delays.data_ptr() if lean else None,
                                              N_ROWS, T, S, EPS, 0, flags, status.data_ptr(), st)
        _lib.check(rc, "simulst_mma_train_fwd_delays")
        rc = lib.simulst_mma_train_bwd_delays(pd.data_ptr(), B16, e.data_ptr(), B16, None, alpha.data_ptr(),
                                              side.data_ptr(), None if lean else ga.data_ptr(), gb.data_ptr(),
                                              gd.data_ptr() if lean else None, gpd.data_ptr(), B16, ge.data_ptr(), B16,
                                              N_ROWS, T, S, EPS, 0, flags, st)
        _lib.check(rc, "simulst_mma_train_bwd_delays")

    peak, _ = measured_peak()
    elems = N_ROWS * T * S
    e_in = 2
    out = {"metric": "mma_pooled_fwd_bwd_elements_per_s", "unit": UNIT,
           "config": f"training shape N={N_ROWS} tgt={T} src={S}, fixed pre-decision ratio {ratio}: p_choose_pooled "
                     f"[N,T,{sp}] bf16 in, pooled gradient out; working set >> L2; 4 launches per step",
           "p_choose_bytes_read_per_element": e_in / ratio, "dense_p_choose_bytes_per_element": e_in}
    for name, lean in (("full_outputs", False), ("latency_loss_mode", True)):
        for _ in range(3):
            step_pooled(lean)
            step_dense(lean)
        torch.cuda.synchronize()
        us_p = _events_us(lambda: step_pooled(lean), 10)
        us_d = _events_us(lambda: step_dense(lean), 10)
        # algorithmic bytes per dense element of the pooled call: pooled p + energy in, beta (+ alpha) out;
        # energy, grad_beta (+ grad_alpha) in, grad_energy + pooled grad out; grid intermediates 4/ratio each
        grid = 4.0 / ratio
        bytes_f = e_in / ratio + e_in + 4 + (0 if lean else 4) + 2 * grid
        bytes_b = e_in / ratio + e_in + 4 + (0 if lean else 4) + e_in + e_in / ratio + 4 * grid
        out[name] = {"pooled_us_per_step": us_p, "dense_kernels_on_expanded_tensor_us_per_step": us_d,
                     "value": elems / (us_p * 1e-6), "speedup_vs_dense_kernels": us_d / us_p,
                     "algorithmic_bytes_per_element": bytes_f + bytes_b,
                     "roofline_frac": elems * (bytes_f + bytes_b) / (us_p * 1e-6) / 1e9 / peak}
    # right-padded batch (lengths ~ U[S/2, S], the masked_batch entry's mask) with the right-padding promise
    for _ in range(3):
        step_pooled(False, True)
        step_dense(False, True)
    torch.cuda.synchronize()
    us_p = _events_us(lambda: step_pooled(False, True), 10)
    us_d = _events_us(lambda: step_dense(False, True), 10)
    out["right_padded_batch"] = {"pooled_us_per_step": us_p, "dense_kernels_on_expanded_tensor_us_per_step": us_d,
                                 "value": elems / (us_p * 1e-6), "speedup_vs_dense_kernels": us_d / us_p}
    out["value"] = out["full_outputs"]["value"]
    return out


def _cpu_best(fn, reps=2):
    """Best wall time of `fn` over `reps` runs after one warm-up, all host cores."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    fn()
    best = float("inf")
    for _ in range(reps):
        c0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - c0)
    return best, torch.get_num_threads()


def bench_ssnt(lib, dev):
    """ssnt_loss forward + backward (criterion/ssnt_loss/ssnt_loss.py:45-151) next to
    the reference function on this box's cores, same inputs."""
    import torch
    from simulst_b200.criterion.ssnt_loss import ssnt_loss
    n, t, s, v = 16, 48, 192, 256
    g = torch.Generator().manual_seed(6100)
    logits = torch.randn(n, t, s, v, generator=g)
    emit = torch.randn(n, t, s, generator=g) - 1.0
    targets = torch.randint(0, v, (n, t), generator=g)
    src_len = torch.randint(s // 2, s + 1, (n,), generator=g)
    tgt_len = torch.randint(t // 2, t + 1, (n,), generator=g)
    src_len[0], tgt_len[0] = s, t
    lp_d = logits.to(dev).log_softmax(-1).requires_grad_()
    em_d = emit.to(dev).requires_grad_()
    args_d = (targets.to(dev), src_len.to(dev), tgt_len.to(dev))

    def step():
        lp_d.grad = None
        em_d.grad = None
        loss, _, _ = ssnt_loss(lp_d, *args_d, emit_logits=em_d, reduction="sum")
        loss.backward()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    us = _events_us(step, 10)
    out = {"metric": "ssnt_loss_fwd_bwd_us", "value": us, "unit": "us",
           "config": f"N={n} T={t} S={s} V={v} fp32, ragged lengths, emit logits, reduction sum; through the Python mirror "
                     "(log_probs gradient [N,T,S,V] included)",
           "lattice_cells_per_s": n * t * s / (us * 1e-6)}
    try:
        from oracle import ref_loader
        if ref_loader.available():
            ref = ref_loader.load_ssnt().ssnt_loss
            lp_c = logits.log_softmax(-1).requires_grad_()
            em_c = emit.clone().requires_grad_()

            def cpu_step():
                lp_c.grad = None
                em_c.grad = None
                loss_c, _, _ = ref(lp_c, targets, src_len, tgt_len, emit_logits=em_c, reduction="sum")
                loss_c.backward()
                return loss_c
            sec, cores = _cpu_best(cpu_step)
            l_ref = float(cpu_step().detach())
            l_got = float(step().detach())
            if abs(l_got - l_ref) > 1e-4 * abs(l_ref):
                raise SystemExit(f"ssnt bench parity check failed: {l_got} vs {l_ref}")
            out["cpu_baseline"] = {"value": sec * 1e6, "unit": "us", "cores": cores, "kind": "reference",
                                   "sample": "full config, best of 2 after 1 warm-up",
                                   "loss_rel_diff": abs(l_got - l_ref) / abs(l_ref)}
    except SystemExit:
        raise
    except Exception as exc:  # pragma: no cover
        out["cpu_baseline"] = {"error": repr(exc)}
    return out


def bench_ctc(lib, dev):
    """CTC best alignment (criterion/best_alignment: the reference's one native
    kernel + its S-iteration Python back-trace) at the CIF criterion's shape: B=64 utterances,
    375 encoder frames, 60-token targets, V=1024.  Baselines: the reference's Python wrapper over
    its own JIT-built CUDA kernel on this GPU (when it builds here), checked equal."""
    import torch
    from simulst_b200.criterion.best_alignment import best_alignment
    b, s, t, v = 64, 375, 60, 1024
    g = torch.Generator().manual_seed(6200)
    lp = torch.randn(s, b, v, generator=g).log_softmax(-1).to(dev)
    targets = torch.randint(1, v, (b, t), generator=g).to(dev)
    in_len = torch.randint(s // 2, s + 1, (b,), generator=g).to(dev)
    tg_len = torch.randint(t // 2, t + 1, (b,), generator=g).to(dev)
    in_len[0], tg_len[0] = s, t
    for _ in range(3):
        states = best_alignment(lp, targets, in_len, tg_len)
    torch.cuda.synchronize()
    us = _events_us(lambda: best_alignment(lp, targets, in_len, tg_len), 10)
    out = {"metric": "ctc_best_alignment_us", "value": us, "unit": "us",
           "config": f"B={b} S={s} T={t} V={v} fp32 log-probs, ragged lengths; one launch incl. the back-trace",
           "frames_per_s": b * s / (us * 1e-6)}
    try:
        from oracle import ref_loader
        if ref_loader.available():
            ext = ref_loader.load_best_alignment_extension()
            ref_call = ref_loader.load_best_alignment_python()
            ref_states = ref_call(ext, lp, targets, in_len, tg_len)
            torch.cuda.synchronize()
            ref_us = _events_us(lambda: ref_call(ext, lp, targets, in_len, tg_len), 3)
            out["reference_gpu"] = {"value": ref_us, "unit": "us", "kind": "reference",
                                    "what": "criterion/best_alignment/__init__.py over the reference's own CUDA kernel "
                                            "(JIT-built here), same GPU",
                                    "states_equal": bool(torch.equal(states, ref_states))}
            if not out["reference_gpu"]["states_equal"]:
                raise SystemExit("ctc bench parity check failed: alignment differs from the reference's")
    except SystemExit:
        raise
    except Exception as exc:  # pragma: no cover
        out["reference_gpu"] = {"error": repr(exc)[:300]}
    return out


def bench_dal(lib, dev):
    """DifferentiableAverageLagging forward + backward on the [N,T] expected delays
    (criterion/mma_criterion.py:172-177), against SimulEval's formulation (restated in
    oracle/latency.py -- the function is not vendored by the reference) on this box's cores."""
    import torch
    from simulst_b200 import ops
    n, t = 64 * 6 * 8, 128          # batch x layers x heads rows, as compute_latency_loss flattens them
    g = torch.Generator().manual_seed(6300)
    delays = (torch.rand(n, t, generator=g) * 1000.0).cumsum(1) / 64.0
    src = torch.randint(512, 1025, (n,), generator=g)
    mask = torch.arange(t)[None, :] >= torch.randint(t // 2, t + 1, (n, 1), generator=g)
    d_d = delays.to(dev).requires_grad_()
    src_d, mask_d = src.to(dev), mask.to(dev)

    def step():
        d_d.grad = None
        out = ops.differentiable_average_lagging(d_d, src_d, None, mask_d)
        out.sum().backward()
        return out
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    us = _events_us(step, 10)
    res = {"metric": "dal_fwd_bwd_us", "value": us, "unit": "us", "config": f"{n} rows x tgt {t}, target padding mask"}
    try:
        from oracle import latency as olat
        d_c = delays.clone().requires_grad_()

        def cpu_step():
            d_c.grad = None
            o = olat.differentiable_average_lagging(d_c, src, None, mask)
            o.sum().backward()
            return o
        sec, cores = _cpu_best(cpu_step)
        diff = float((step().detach().cpu() - cpu_step().detach()).abs().max())
        res["cpu_baseline"] = {"value": sec * 1e6, "unit": "us", "cores": cores, "kind": "port",
                               "sample": "full config, best of 2 after 1 warm-up", "max_abs_diff": diff}
    except Exception as exc:  # pragma: no cover
        res["cpu_baseline"] = {"error": repr(exc)[:300]}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="row chunks of the host-buffer pipeline")
    ap.add_argument("--e2e-streams", type=int, default=2, help="compute streams of the host-buffer pipeline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun when started as a plain process
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup)]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
