#!/usr/bin/env python
"""bench.py — collocation points/sec of one NestedLoRA loss+grad step (2D hydrogen, L=16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = fused forward (K1) + Gram/loss (K2) + dF (K3) + MLP backward (K4) on P synthetic
collocation points PER GPU (weak scaling), plus the two all-reduces when N > 1.
  value : whole-job points/s with x already resident in HBM (CUDA events, max over ranks);
  e2e   : same through the public API `NestedLoRA.compute_loss_operator` + backward with x in
          pinned HOST memory (H2D inside the timed region, loss.item() D2H);
  roofline     : the layer-0 forward GEMM (tcgen05), timed with CUDA events inside the timed steps;
  cpu_baseline : the reference's own CPU PyTorch path (baseline/_ref, unmodified) or, if that copy
                 is absent, the numpy oracle port, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "collocation points/sec (loss+grad step), 2D hydrogen L=16"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=131072, help="collocation points per GPU per step")
    ap.add_argument("--neigs", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="f16x3", choices=["f16x3", "bf16x3", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-BASELINE-config block")
    return ap.parse_args()


def workload_config(P, L, world):
    """`config` of the JSON line; the reference arm prints the SAME dict (it times a bounded sample of this workload)."""
    return {"workload": f"2D hydrogen (H=-Lap-1/r, scale 100), joint nesting, L={L}, M_ff=1024, "
                        f"{P} Gaussian(sigma=16) collocation points per GPU per step, exact Laplacian "
                        f"(laplacian_eps=0), loss+grad", "points_per_gpu": P, "neigs": L,
            "parallelism": f"dp{world} over points", "laplacian": "exact",
            "l2": "512 MB flush write between timed steps"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU during the timed region (NVML, 100 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "app_clocks": 0x2}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, budget_s, steps=None, warmup=1):
    """Times the reference's CPU implementation of the step on the host cores.
    kind 'reference' = unmodified jongharyu/neural-svd from baseline/_ref (or /root/reference),
    script-default finite-difference Laplacian (eps=0.01, hydrogen.sh:20) AND exact mode;
    kind 'port' = oracle/nsvd_oracle.py (numpy, forward-mode exact Laplacian)."""
    import numpy as np
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import nsvd_oracle as O
    from oracle import ref_bootstrap as RB
    root = RB.find_reference()
    out = {"cores": cores}
    if root is not None:
        ref = RB.import_reference(root)
        B = 512
        x = (cfg.sampling_scale * torch.randn((B, 1, cfg.ndim))).reshape(B, -1)
        res = {}
        for tag, eps in (("fd", 0.01), ("exact", 0.0)):
            method, operator, importance, _ = RB.build_reference_problem(ref, cfg, 0, eps)
            ts = []
            t_all = time.perf_counter()
            n = 0
            while True:
                t0 = time.perf_counter()
                method.zero_grad()
                loss, _ = method.compute_loss_operator(operator, x, importance=importance)
                loss.backward()
                float(loss.detach())
                dt = time.perf_counter() - t0
                if n >= warmup:
                    ts.append(dt)
                n += 1
                if steps is not None and len(ts) >= steps:
                    break
                if steps is None and (time.perf_counter() - t_all > budget_s / 2 and len(ts) >= 2):
                    break
            res[tag] = B / statistics.median(ts)
        out.update(kind="reference", value=res["exact"], unit=UNIT, exact_value=res["exact"], fd_value=res["fd"],
                   sample_points=B,
                   sample=f"unmodified reference (torch {torch.__version__} CPU, {cores} threads), hydrogen L={cfg.neigs}, "
                          f"B={B} points/step, loss+grad; value = exact_value = autograd exact Laplacian (laplacian_eps=0: the "
                          f"operator the GPU arm evaluates, and the parity oracle), fd_value = finite-difference Laplacian "
                          f"eps=0.01 (the scripts' default, hydrogen.sh:20)")
        return out
    B = 2048
    params = O.init_params_like_reference(cfg, 0)
    x = (cfg.sampling_scale * np.random.RandomState(0).randn(B, cfg.ndim)).astype(np.float32)
    ts, t_all = [], time.perf_counter()
    while True:
        t0 = time.perf_counter()
        O.train_step(x, params, cfg)
        ts.append(time.perf_counter() - t0)
        if (steps is not None and len(ts) >= steps) or (steps is None and time.perf_counter() - t_all > budget_s and len(ts) >= 2):
            break
    out.update(kind="port", value=B / statistics.median(ts), unit=UNIT, sample_points=B,
               sample=f"numpy oracle port (fp32, forward-mode exact Laplacian, {cores} BLAS threads), hydrogen "
                      f"L={cfg.neigs}, B={B} points/step")
    return out


def run_reference_arm(args):
    """The reference's own CPU implementation of the SAME workload (same operator: exact Laplacian, laplacian_eps=0),
    each step a bounded sample (B=512 points) of it; the script-default finite-difference mode is reported next to it
    (cpu_baseline.fd_value)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import nsvd_oracle as O
    cfg = O.PathConfig.hydrogen(neigs=args.neigs)
    t0 = time.perf_counter()
    r = cpu_reference_run(cfg, budget_s=60.0, steps=max(args.steps, 2), warmup=max(args.warmup, 1))
    v = r["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["sample_points"] / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.points, args.neigs, args.gpus),
            "cpu_baseline": r, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# problem builders (the product API only: no oracle, no test code) and per-config timing
# ------------------------------------------------------------------------------------------
def script_args(kind, neigs, laplacian_eps=0.0, ndim=2):
    """Hyper-parameters of scripts/exps/pde/{hydrogen,oscillator}.sh with BASELINE's L."""
    from types import SimpleNamespace
    base = dict(problem="sch", ndim=ndim, neigs=neigs, charge=1.0, laplacian_eps=laplacian_eps, lim=50.0, use_fourier_feature=True,
                fourier_deterministic=False, fourier_append_raw=False, mlp_hidden_dims="128,128,128",
                nonlinearity="softplus", parallel=True, apply_boundary=False, boundary_mode="dir_box_sqrt",
                hard_mul_const=1.0)
    if kind == "hydrogen":          # hydrogen.sh:11-65
        base.update(potential_type="hydrogen", operator_scale=100.0, operator_shift=0.0, fourier_mapping_size=1024,
                    fourier_scale=0.1, apply_exp_mask=False, exp_mask_init_scale=100.0, sampling_scale=16.0)
    else:                           # oscillator.sh:11-67
        base.update(potential_type="harmonic_oscillator", operator_scale=1.0, operator_shift=16.0,
                    fourier_mapping_size=256, fourier_scale=1.0, apply_exp_mask=True, exp_mask_init_scale=10.0,
                    sampling_scale=4.0)
    return SimpleNamespace(**base)


def make_problem(N, kind, neigs, sequential, dev, seed=0, laplacian_eps=0.0, ndim=2):
    import torch
    cfg = script_args(kind, neigs, laplacian_eps, ndim)
    torch.manual_seed(seed)
    operator, _gt = N.get_problem(cfg)
    model = N.get_wavefunctions(cfg)
    method = N.NestedLoRA(model=model, neigs=neigs, step=1, sort=False, sequential=sequential).to(dev)
    return cfg, method, operator, N.GaussianImportance(cfg.sampling_scale, cfg.ndim)


def time_events(fn, steps, warmup, sync):
    """ms per call of fn() over `steps` calls after `warmup`, CUDA events on the current stream."""
    import torch
    for _ in range(warmup):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / steps


def small_config(N, kind, B, neigs, sequential, dev, steps=50, warmup=10, laplacian_eps=0.0):
    """eager and CUDA-graph step time of one of the reference's own small-batch configurations (1 GPU)."""
    import torch
    cfg, method, operator, importance = make_problem(N, kind, neigs, sequential, dev, laplacian_eps=laplacian_eps)
    g = torch.Generator().manual_seed(7)
    xs = [(cfg.sampling_scale * torch.randn((B, 1, 2), generator=g)).reshape(B, 2).to(dev) for _ in range(4)]
    it = [0]

    def eager():
        method.zero_grad(set_to_none=True)
        loss, _ = method.compute_loss_operator(operator, xs[it[0] % 4], importance=importance)
        loss.backward()
        it[0] += 1

    sync = lambda: torch.cuda.synchronize(dev)
    ms_eager = time_events(eager, steps, warmup, sync)
    method.zero_grad(set_to_none=True)
    gstep = N.GraphedOperatorStep(method, operator, importance, B)
    ms_graph = time_events(lambda: gstep(xs[0]), steps, warmup, sync)
    return {"points": B, "neigs": neigs, "nesting": "sequential" if sequential else "joint", "problem": kind,
            "laplacian": f"finite differences, eps={laplacian_eps}" if laplacian_eps > 0 else "exact",
            "ms_per_step_eager": ms_eager, "ms_per_step_graphed": ms_graph,
            "points_per_s_eager": B / (ms_eager * 1e-3), "points_per_s_graphed": B / (ms_graph * 1e-3)}


def cdk_config(N, dev, B=4096, L=512, steps=20, warmup=5):
    """BASELINE config 5: CDK NestedLoRA loss forward + backward on (B, L) embeddings (+1 constant mode), joint."""
    import torch
    g = torch.Generator().manual_seed(10)
    f = torch.randn(B, L, generator=g).to(dev).requires_grad_()
    gg = torch.randn(B, L, generator=g).to(dev).requires_grad_()
    method = N.NestedLoRAForCDK(model=None, neigs=L, step=1, sequential=False, set_first_mode_const=True)
    out = {}
    for tag, diag in (("with_diagnostics", True), ("without_diagnostics", False)):
        method.diagnostics = diag

        def step():
            f.grad = gg.grad = None
            loss = method.compute_loss(f, gg)[0]
            loss.backward()

        ms = time_events(step, steps, warmup, lambda: torch.cuda.synchronize(dev))
        out[tag] = {"ms_per_step": ms, "pairs_per_s": B / (ms * 1e-3)}
    out.update(rows=B, feature_dim=L, what="NestedLoRAForCDK.compute_loss + backward (loss only, no encoder)")
    # the whole CDK step of main_sketchy.py:176-186: two 512 -> 8192 -> 512 towers (dense layers on the library's own
    # tcgen05 GEMM kernels) -> l2_ball -> fused CDK loss -> backward into the tower weights; next to it the same
    # towers on torch / cuBLAS fp32 (TF32 off: same arithmetic contract) feeding the same loss kernels
    net = N.get_sketchy_encoder().to(dev)
    x, y = torch.randn(B, 512, generator=g).to(dev), torch.randn(B, 512, generator=g).to(dev)
    m2 = N.NestedLoRAForCDK(model=net, neigs=L, step=1, sequential=False, set_first_mode_const=True).to(dev)
    m2.diagnostics = False

    def enc_step():
        m2.zero_grad(set_to_none=True)
        _, fx, _, fy = m2(x, y)
        m2.compute_loss(fx, fy)[0].backward()

    ms = time_events(enc_step, steps, warmup, lambda: torch.cuda.synchronize(dev))
    flop = 2 * 3 * 2 * B * (512 * 8192 + 8192 * 512)          # 2 towers x (fwd + dgrad + wgrad) x 2 B K N per layer
    out["with_encoder"] = {"ms_per_step": ms, "pairs_per_s": B / (ms * 1e-3), "encoder_gflop": flop / 1e9,
                           "encoder_tflops_algorithmic": flop / (ms * 1e-3) / 1e12,
                           "what": "HeteroNetwork towers (TCLinear: nsvd_linear_fwd/bwd) + CDK loss + backward"}
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    towers = [torch.nn.Sequential(torch.nn.Linear(512, 8192), torch.nn.LeakyReLU(0.2), torch.nn.Linear(8192, 512)).to(dev)
              for _ in range(2)]

    def cublas_step():
        for t in towers:
            t.zero_grad(set_to_none=True)
        fx, fy = N.normalize(towers[0](x), 4.0, "l2_ball"), N.normalize(towers[1](y), 4.0, "l2_ball")
        m2.compute_loss(fx, fy)[0].backward()

    ms_c = time_events(cublas_step, steps, warmup, lambda: torch.cuda.synchronize(dev))
    torch.backends.cuda.matmul.allow_tf32 = tf32
    out["with_encoder"]["cublas_fp32_ms_per_step"] = ms_c
    return out


def strong_scaling_config(N, dev, dp, world, rank, total_points=1 << 20, neigs=64, steps=3, warmup=2):
    """BASELINE config 4: 2^20 hydrogen collocation points per step, L=64, joint, sharded over the ranks (STRONG
    scaling: total work fixed), including both all-reduces."""
    import torch
    import torch.distributed as dist
    P = total_points // world
    cfg, method, operator, importance = make_problem(N, "hydrogen", neigs, False, dev)
    method.data_parallel = dp
    g = torch.Generator().manual_seed(300 + rank)
    x = (cfg.sampling_scale * torch.randn((P, 1, 2), generator=g)).reshape(P, 2).to(dev)

    def step():
        method.zero_grad(set_to_none=True)
        loss, _ = method.compute_loss_operator(operator, x, importance=importance)
        loss.backward()

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ms = time_events(step, steps, warmup, sync)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    L, K0 = neigs, 2 * cfg.fourier_mapping_size
    flop_pt = L * (2 * 4 * (K0 * 128 + 2 * 128 * 128 + 128) + 2 * (K0 * 128 + 4 * 128 * 128 + 2 * 128))
    return {"total_points": total_points, "points_per_gpu": P, "neigs": neigs, "n_gpus": world, "scaling": "strong",
            "ms_per_step": ms, "points_per_s": total_points / (ms * 1e-3),
            "tflops_algorithmic": flop_pt * total_points / (ms * 1e-3) / 1e12,
            "grad_allreduce_mbytes": sum(p.numel() for p in method.parameters() if p.requires_grad) * 4 / 1e6}

# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def hbm_kernel_probe(N, dev, hbm_peak_gbps, B=1 << 20):
    """K2 (nsvd_gram_reduce, 8 B L algorithmic bytes) and K3 (nsvd_loss_dF, 12 B L) alone at 2^20 points - large enough to
    be HBM-bound instead of launch-latency sized as at the headline batch - against the measured copy peak.
    Direct C-ABI calls back to back on the current stream, CUDA events, inputs (134 / 537 MB) larger than the L2."""
    import ctypes as C
    import torch
    from neural_svd_b200 import _lib
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    out = {"points": B, "peak_GBps": hbm_peak_gbps}
    for L in (16, 64):
        F, TF = torch.randn(B, L, device=dev), torch.randn(B, L, device=dev)
        v = torch.ones(L, device=dev)
        terms = torch.empty(2 * L * L + 5, device=dev)
        coef = torch.randn(2 * L * L + 1, device=dev)
        dF = torch.empty_like(F)
        part = torch.empty(lib.nsvd_gram_partials_bytes(B, L), dtype=torch.uint8, device=dev)

        def k2():
            _lib.check(lib.nsvd_gram_reduce(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), B, L, B // 2, _lib.ptr(terms),
                                            _lib.ptr(part), st), "nsvd_gram_reduce")

        def k3():
            _lib.check(lib.nsvd_loss_dF(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), _lib.ptr(coef), None, B, L, B // 2, B,
                                        _lib.ptr(dF), st), "nsvd_loss_dF")

        for name, fn, nbytes in (("gram_reduce", k2, 8.0 * B * L), ("loss_dF", k3, 12.0 * B * L)):
            ms = time_events(fn, 20, 3, lambda: torch.cuda.synchronize(dev))
            gbps = nbytes / (ms * 1e-3) / 1e9
            out[f"{name}_L{L}"] = {"us": ms * 1e3, "algorithmic_mbytes": nbytes / 1e6, "achieved_GBps": gbps,
                                   "hbm_frac": gbps / hbm_peak_gbps}
        del F, TF, dF, part
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import neural_svd_b200 as N
    from neural_svd_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N>1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    _lib.check(lib.nsvd_device_ok(local), "nsvd_device_ok")
    dp = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dp = N.PointParallel()

    N.set_engine(args.engine)
    # the HBM-bound kernels timed ALONE, before the tensor-core loops push the chip onto its power cap (the SM clock stays
    # low for a while afterwards, and these short kernels are issue / latency sensitive): burst conditions, stated as such
    hbm_probe = None
    if world == 1 and not args.no_configs:
        hbm_probe = hbm_kernel_probe(N, dev, peaks()["hbm"])
        hbm_probe["when"] = "timed alone at the start of the run (before the step loops; SM clock not yet power-capped)"
    cfg, method, operator, importance = make_problem(N, "hydrogen", args.neigs, False, dev)
    method.data_parallel = dp
    P = args.points
    g = torch.Generator().manual_seed(100 + rank)
    xs_host = [(cfg.sampling_scale * torch.randn((P, 1, 2), generator=g)).reshape(P, 2).pin_memory()
               for _ in range(4)]
    xs_dev = [x.to(dev) for x in xs_host]
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step(x):
        method.zero_grad(set_to_none=True)
        loss, _ = method.compute_loss_operator(operator, x, importance=importance)
        loss.backward()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(xs_dev[i % 4])
    barrier()

    # ---- device-resident timing: K steps, L2 flushed between steps, CUDA events on the launch stream
    sampler = ClockSampler(local)
    sampler.start()
    lib.nsvd_profile_enable(1)
    _lib.profile_read(reset=True)
    launches0 = lib.nsvd_launch_count()
    evs = []
    barrier()
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(xs_dev[i % 4])
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = lib.nsvd_launch_count() - launches0
    lib.nsvd_profile_enable(0)
    prof = _lib.profile_read(reset=True)
    sampler.stop_flag = True
    sampler.join()
    ms = [a.elapsed_time(b) for a, b in evs]
    t_dev = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev)
    ms_per_step = total_ms / args.steps
    value = world * P / (ms_per_step * 1e-3)

    # ---- end to end: host x (pinned) -> H2D -> step -> loss.item(); wall clock, max over ranks
    for i in range(2):
        step(xs_host[i % 4]).detach().item()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(xs_host[i % 4]).detach().item()     # .to(device) inside compute_loss_operator; .item() reads the loss back
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * P * args.steps / float(t_e2e)

    # ---- second column: sampler + step + fused RMSprop/EMA/cosine update, all on the device
    fopt = N.FusedRMSpropEMA(method.parameters(), lr=1e-4, alpha=0.999, eps=1e-10, ema_decay=0.995, num_iters=10 ** 6)

    def full_step(i):
        x = N.sample_gaussian(P, cfg.sampling_scale, seed=1234 + rank, offset=i * P, device=dev)
        loss = step(x)
        fopt.step()
        return loss

    for i in range(2):
        full_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        full_step(2 + i)
    f1.record()
    barrier()
    t_full = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_full, op=dist.ReduceOp.MAX)
    full_ms = float(t_full) / args.steps

    # ---- third column: the same step replayed as ONE CUDA graph (GraphedOperatorStep; single GPU only)
    graphed = None
    gstep = None
    if world == 1:
        method.zero_grad(set_to_none=True)
        try:
            gstep = N.GraphedOperatorStep(method, operator, importance, P)     # owns a second set of scratch buffers
        except torch.OutOfMemoryError:
            gstep = None
            graphed = {"skipped": "a second set of scratch buffers does not fit next to the eager one at this size"}
        if gstep is not None:
            for i in range(2):
                gstep(xs_dev[i % 4])
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for i in range(args.steps):
                gstep(xs_dev[i % 4])
            g1.record()
            barrier()
            gms = g0.elapsed_time(g1) / args.steps
            graphed = {"value": P / (gms * 1e-3), "unit": UNIT, "ms_per_step": gms,
                       "what": "loss+grad step replayed as one CUDA graph (no L2 flush between steps)"}

    # ---- the other BASELINE.json configurations (each at the same parity bar in tests/; here: their step times)
    configs = {}
    if not args.no_configs:
        del gstep
        xs_dev.clear()
        method.zero_grad(set_to_none=True)
        method.__dict__.pop("_nsvd_scratch", None)
        del flush
        torch.cuda.empty_cache()
        if world == 1:
            configs["config1_hydrogen_b128_seq_L16"] = small_config(N, "hydrogen", 128, 16, True, dev)
            configs["config2_oscillator_b512_jnt_L16"] = small_config(N, "oscillator", 512, 16, False, dev)
            configs["config3_hydrogen_b512_jnt_L16"] = small_config(N, "hydrogen", 512, 16, False, dev)
            configs["config5_cdk_b4096_L512"] = cdk_config(N, dev)
            # the scripts' own Laplacian mode (laplacian_eps = 0.01, hydrogen.sh:20): second, value-only pass over the four
            # shifted point sets; same batch as config 3 and the headline workload
            configs["config3_fd_eps0p01"] = small_config(N, "hydrogen", 512, 16, False, dev, laplacian_eps=0.01)
            cfd, mfd, ofd, ifd = make_problem(N, "hydrogen", args.neigs, False, dev, laplacian_eps=0.01)
            xfd = (cfd.sampling_scale * torch.randn((P, 1, 2))).reshape(P, 2).to(dev)

            def stepfd():
                mfd.zero_grad(set_to_none=True)
                loss, _ = mfd.compute_loss_operator(ofd, xfd, importance=ifd)
                loss.backward()

            msfd = time_events(stepfd, 5, 2, lambda: torch.cuda.synchronize(dev))
            configs["headline_workload_fd_eps0p01"] = {"points": P, "ms_per_step": msfd, "points_per_s": P / (msfd * 1e-3)}
            mfd.__dict__.pop("_nsvd_scratch", None)
            del mfd, xfd
            torch.cuda.empty_cache()
            # reference-grade CUDA-core engine on the headline workload, once (validation engine, not the product path)
            N.set_engine("fp32")
            c32, m32, o32, i32 = make_problem(N, "hydrogen", args.neigs, False, dev)
            x32 = (c32.sampling_scale * torch.randn((P, 1, 2))).reshape(P, 2).to(dev)

            def step32():
                m32.zero_grad(set_to_none=True)
                loss, _ = m32.compute_loss_operator(o32, x32, importance=i32)
                loss.backward()

            ms32 = time_events(step32, 2, 1, lambda: torch.cuda.synchronize(dev))
            configs["fp32_engine_headline_workload"] = {"points": P, "ms_per_step": ms32, "points_per_s": P / (ms32 * 1e-3)}
            del m32, x32
            # ndim = 3 hydrogen (pde/problems.py:62-68): five forward-mode streams on the CUDA-core engine
            c3, m3, o3, i3 = make_problem(N, "hydrogen", args.neigs, False, dev, ndim=3)
            x3 = (c3.sampling_scale * torch.randn((16384, 1, 3))).reshape(16384, 3).to(dev)

            def step3d():
                m3.zero_grad(set_to_none=True)
                loss, _ = m3.compute_loss_operator(o3, x3, importance=i3)
                loss.backward()

            ms3 = time_events(step3d, 3, 1, lambda: torch.cuda.synchronize(dev))
            configs["hydrogen_ndim3_fp32_engine"] = {"points": 16384, "neigs": args.neigs, "ms_per_step": ms3,
                                                     "points_per_s": 16384 / (ms3 * 1e-3)}
            del m3, x3
            N.set_engine(args.engine)
            torch.cuda.empty_cache()
        configs["config4_hydrogen_2p20_L64_strong"] = strong_scaling_config(N, dev, dp, world, rank)

    if rank == 0:
        pk = peaks()
        L, K0 = args.neigs, 2 * cfg.fourier_mapping_size
        l0_ms, l0_n = prof["l0_fwd"]
        flops_l0 = 2.0 * 4 * K0 * 128 * L * P * args.steps          # algorithmic, all launches of the timed region
        roof = None
        if l0_n:
            ach = flops_l0 / (l0_ms * 1e-3) / 1e12
            # DRAM traffic of this kernel from the committed ncu capture of the final build (profiles/r2g_ncu_full.csv:
            # 4.551 GB read + 2.531 GB written by one launch over 65536 points; other captures of the same kernel read
            # 2.9 - 6.7 GB: how much of a 32 MB Phi group survives in the L2 varies), scaled to the points of a launch here
            pts_per_launch = P * args.steps / l0_n
            traffic = (4.550687e9 + 2.530890e9) / 65536 * pts_per_launch
            roof = {"kernel": "big2s_gemm_kernel<K-major, L0FwdEpi> (layer-0 4-stream forward GEMM, tcgen05 cta_group::2, fp16 hi/lo planes x 3 products)",
                    "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sust"], "traffic": traffic,
                    "traffic_note": "bytes per launch = ncu dram read+write per point (108.1 KB, profiles/r2g_ncu_full.csv) x "
                                    "points per launch; algorithmic bytes are 8 KB (Phi) + 32 KB (3 derivative streams + the saved value "
                                    "stream out) per point + 67 MB of folded weights per 4096-point group: the Phi group is re-read "
                                    "(L2 thrash next to the output streams); the kernel is tensor-bound (pipe 96 % active, DRAM 24 %)",
                    "peak_source": pk["src"] + " bf16 sustained",
                    "issued_frac": 3 * ach / pk["tf_sust"], "launches": l0_n, "avg_launch_ms": l0_ms / l0_n,
                    "note": "achieved = ALGORITHMIC fp32-equivalent FLOPs; every MAC is issued as 3 fp16 MMAs "
                            "(hi*hi + hi*lo + lo*hi), so tensor-pipe issue rate = issued_frac of the 16-bit dense peak"}
        gram_ms, gram_n = prof["gram_reduce"]
        df_ms, df_n = prof["loss_dF"]
        kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()}
        if gram_n:
            kernels["gram_reduce"]["achieved_GBps"] = 8.0 * P * L * args.steps / (gram_ms * 1e-3) / 1e9
            kernels["gram_reduce"]["hbm_frac"] = kernels["gram_reduce"]["achieved_GBps"] / pk["hbm"]
        if df_n:
            kernels["loss_dF"]["achieved_GBps"] = 12.0 * P * L * args.steps / (df_ms * 1e-3) / 1e9
            kernels["loss_dF"]["hbm_frac"] = kernels["loss_dF"]["achieved_GBps"] / pk["hbm"]
        if hbm_probe is not None:
            # the HBM-bound kernels alone at a size where the HBM, not the launch latency, bounds them (the two entries
            # above are timed inside the step at the headline batch: 17 - 25 MB, latency sized)
            kernels["hbm_probe_2p20_points"] = hbm_probe
        flop_pt = L * (2 * 4 * (K0 * 128 + 2 * 128 * 128 + 128) + 2 * (K0 * 128 + 4 * 128 * 128 + 2 * 128))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16x3 (fp32 operands as fp16 hi/lo planes, 3 tensor-core products, fp32 accumulate)" if args.engine == "f16x3" else "f32",
                "data": "synthetic",
                "config": workload_config(P, L, world), "engine": args.engine,
                "algorithmic_mflop_per_point": flop_pt / 1e6,
                "clocks": sampler.result(), "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": P * 2 * 4 * world,
                        "d2h_bytes_per_step": 4 * world},
                "with_optimizer": {"value": world * P / (full_ms * 1e-3), "unit": UNIT, "ms_per_step": full_ms,
                                   "what": "device sampler + loss+grad + fused RMSprop/EMA/cosine update (no L2 flush)"},
                "graphed": graphed, "roofline": roof, "kernels": kernels,
                "step_tflops_algorithmic": flop_pt * value / 1e12, "configs": configs}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import nsvd_oracle as O        # cpu_baseline leg only
            cb = line["cpu_baseline"] = cpu_reference_run(O.PathConfig.hydrogen(neigs=args.neigs), args.cpu_seconds)
            c3 = configs.get("config3_hydrogen_b512_jnt_L16")
            if c3 and cb.get("kind") == "reference":
                # same problem, same batch (B=512), same operator (exact Laplacian): GPU step vs the reference on the host
                c3["cpu_reference_exact_points_per_s"] = cb["exact_value"]
                c3["cpu_reference_fd_points_per_s"] = cb["fd_value"]
                c3["speedup_vs_cpu_exact_same_batch"] = c3["points_per_s_graphed"] / cb["exact_value"]
                cfd3 = configs.get("config3_fd_eps0p01")
                if cfd3:    # finite differences on both sides, same batch
                    cfd3["cpu_reference_fd_points_per_s"] = cb["fd_value"]
                    cfd3["speedup_vs_cpu_fd_same_batch"] = cfd3["points_per_s_graphed"] / cb["fd_value"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
