#!/usr/bin/env python
"""Headline benchmark: fp64 Golub-Kahan iterations/s on parallel-beam CT (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx 2048 --views 720]

One "step" = one Golub-Kahan iteration (the reference's golub_kahan_update, trips/utilities/decompositions.py:230-255):
v = A^T u - beta v_prev, alpha = ||v||, v /= alpha, u = A v - alpha u_prev, beta = ||u||, u /= beta, with U and V retained.
Workload at N = 1: configs[3]'s geometry (2048^2 pixels, 720 angles, 2896 detector bins, nnz 3.85e9).  Default layout
'implicit': both projectors are matrix-free (every matrix entry re-evaluated in the kernel, bit-identical to the stored
layouts '--layout sell' / '--layout csr').  For N > 1 the problem is fixed (strong scaling): u-space is sharded by
projection angle, v-space by image band, and the two vectors a step needs everywhere are exchanged by peer-to-peer
stores from the kernel epilogues (trips-py_b200/dist.py).

JSON keys beyond the base contract:
  value      it/s with every input resident in HBM (CUDA events around exactly K steps, max over ranks)
  e2e        it/s through the reference-signature call golub_kahan_update(A, U, S, V) with HOST (NumPy) U, S, V:
             every step copies u_k and v_{k-1} host->device and the new u, v device->host inside the timed region
  roofline   dominant kernel, against the bound that really limits it: `bound: "fp64-issue"` for the matrix-free
             projectors (fp64 instructions per launch / launch time / the DFMA-chain peak measured in this run), with
             the SURVEY 8(d) streaming-equivalent figure as the secondary field `hbm_equivalent`
  parity     bit-for-bit checks of the benchmarked operator against scipy / the oracle, outside the timed region
  secondary  the same step on the stored SELL-32-4 layout (the north_star's CSR SpMV, true HBM bound) and its
             fp32-storage / fp64-accumulate variant, each with its own roofline
  cpu_baseline  the reference's own golub_kahan_update (oracle/_ref) on a bounded sample, on the host cores
--impl reference times that CPU path as its own arm (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "GK iters/sec (fp64) 2048^2 CT"
UNIT = "it/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--views", type=int, default=720)
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-sample-views", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the bit-for-bit checks (they cost ~30 s of host time)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the stored-SELL / fp32-storage records")
    ap.add_argument("--f32-storage", action="store_true", help="fp32-storage / fp64-accumulate variant as the primary record")
    ap.add_argument("--order", default="sequential", choices=["sequential", "tree"],
                    help="SpMV row-sum order: 'sequential' = scipy's (bit-identical to the reference, the parity build); "
                         "'tree' = per-row tree reduction (reported separately)")
    ap.add_argument("--variant", type=int, default=0, help="kernel tuning knob (tb200_spmv_set_variant)")
    ap.add_argument("--layout", default="auto", choices=["auto", "implicit", "sell", "csr"],
                    help="device layout of A / A^T: 'implicit' = matrix-free projectors, 'sell' = stored row-interleaved "
                         "CSR (SELL-32-4), 'csr' = plain CSR; auto = implicit for the fp64 sequential order, sell for "
                         "fp32 storage, csr for tree.  All give bit-identical results in the sequential order")
    ap.add_argument("--exchange", default=None, choices=[None, "p2p", "nccl"],
                    help="N > 1: how the two exchanged vectors travel (default: peer-to-peer stores, NCCL as the transport fallback)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


class ClockSampler:
    """SM clock / throttle reasons / power sampled DURING the timed region through NVML (2 ms period: the timed region
    of 20 steps is ~0.2 s, which a 50 ms nvidia-smi poll can miss entirely)."""

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.t0 = self.t1 = None
        self.h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:  # noqa: BLE001
            self.h = None

    def _loop(self):
        nv, h = self.nv, self.h
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:  # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                self.rows.append((time.time(), sm, reasons, pw))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=1.0)
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "samples_in_timed_region": 0,
                    "power_w_max": None, "source": "nvml unavailable"}
        nv = self.nv
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        rows = inside if inside else self.rows
        sm = sorted(r[1] for r in rows)
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            mx = None
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        reasons = sorted({name for r in rows for name, b in bits.items() if r[2] & b})
        return {"sm_mhz": int(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.rows), "samples_in_timed_region": len(inside),
                "power_w_max": max((r[3] for r in rows), default=None), "source": "nvml, 2 ms period"}


# ---------------------------------------------------------------------------------------------------- CPU arm

def cpu_gk(A_s, b_s, steps, warmup, n_full_rows, nnz_full, use_reference):
    """The reference's golub_kahan_update (oracle/_ref via ref_loader when `use_reference`, else the oracle port) on
    the sample matrix A_s (16 of 720 angles of the FULL 2048^2 image: model-space vectors have their full length,
    data-space vectors 1/45 of it).  Returns the measured sample step and its extrapolation to the full problem:
        t_full = t_spmv * (nnz_full / nnz_s) + t_rest * (n + m_full) / (n + m_s)
    t_spmv = the two sparse products, timed separately on the same vectors; t_rest = everything else in the call
    (norms, scalings, the np.hstack re-copies of U and V, which move (n + m) * k * 8 bytes at step k)."""
    import numpy as np

    if use_reference:
        import ref_loader

        fn, kind = ref_loader.load().decompositions.golub_kahan_update, "reference"
    else:
        import trips_oracle as O

        fn, kind = O.golub_kahan_update, "port"
    U = b_s.reshape(-1, 1) / np.linalg.norm(b_s)
    S, V = np.empty(1), np.empty((A_s.shape[1], 1))
    for _ in range(max(warmup, 1)):
        U, S, V = fn(A_s, U, S, V)
    t0 = time.perf_counter()
    for _ in range(steps):
        U, S, V = fn(A_s, U, S, V)
    t_step = (time.perf_counter() - t0) / steps
    u, v = np.ascontiguousarray(U[:, -1]), np.ascontiguousarray(V[:, -1])
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        A_s.T @ u
        A_s @ v
    t_spmv = (time.perf_counter() - t0) / reps
    t_rest = max(t_step - t_spmv, 0.0)
    m_s, n = A_s.shape
    t_full = t_spmv * (nnz_full / A_s.nnz) + t_rest * (n + n_full_rows) / (n + m_s)
    return {"kind": kind, "t_step_sample_s": t_step, "t_spmv_sample_s": t_spmv, "t_rest_sample_s": t_rest,
            "t_full_s": t_full, "scale_nnz": nnz_full / A_s.nnz, "scale_rows": n_full_rows / m_s,
            "basis_columns_during_timing": int(V.shape[1]),
            "formula": "t_full = t_spmv*(nnz_full/nnz_s) + t_rest*(n+m_full)/(n+m_s)"}


def reference_arm(args, workload, base_config):
    """bench.py --impl reference: the unmodified reference's golub_kahan_update (oracle/_ref) on the host cores."""
    import numpy as np
    import trips_oracle as O
    import ref_loader

    nx, views = args.nx, args.views
    n_det = O.ct_num_detectors(nx)
    sub = np.linspace(0, views, args.cpu_sample_views, endpoint=False).astype(int)
    # the sample matrix comes from the oracle's NumPy builder (bit-identical to the device builder:
    # tests/test_gpu_kernels.py::test_ct_builder_is_bit_identical_to_the_numpy_statement); no GPU code is touched
    A_s = O.ct_matrix(nx, O.ct_angles(views)[sub])
    b_s = A_s @ O.shepp_logan(nx).reshape(-1)
    nnz_full = A_s.nnz * views / len(sub)
    r = cpu_gk(A_s, b_s, args.steps, args.warmup, views * n_det, nnz_full, ref_loader.available())
    rate = 1.0 / r["t_full_s"]
    sample = (f"{len(sub)} of {views} angles of the full {nx}^2 image (nnz {A_s.nnz:.3e} of ~{nnz_full:.3e}); one step = one "
              f"golub_kahan_update on the sample ({r['t_step_sample_s']:.3f} s); value = 1 / t_full, {r['formula']}")
    try:
        from threadpoolctl import threadpool_info

        blas_threads = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
    except Exception:  # noqa: BLE001
        blas_threads = None
    cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": r["kind"], "sample": sample, "host_cpus": os.cpu_count(),
           "blas_threads": blas_threads, "extrapolation": r,
           "note": "scipy's csr/csc matvec (95 % of the step) is single-threaded; the BLAS norms use the OpenBLAS pool"}
    line = {"metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            # ms_per_step = the step that was really timed (one sample step), so steps * ms_per_step is this run's
            # timed region; `value` is the full-size rate it extrapolates to (cpu_baseline.extrapolation)
            "ms_per_step": r["t_step_sample_s"] * 1e3, "ms_per_full_size_step": r["t_full_s"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": base_config,
            "cpu_baseline": cpu,
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------- GPU arm

def measure_fp64_peak(torch, dev):
    """Thread-level fp64 instructions/s of the DFMA-chain microbenchmark (csrc/peaks.cu), best and mean of 8 launches."""
    from trips_b200 import _lib

    L = _lib.lib()
    iters = 20000
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        _lib.check(L.tb200_fp64_peak_run(iters, sink.data_ptr(), st))
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for e0, e1 in evs:
        e0.record()
        _lib.check(L.tb200_fp64_peak_run(iters, sink.data_ptr(), st))
        e1.record()
    torch.cuda.synchronize()
    ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    n_instr = int(L.tb200_fp64_peak_instructions(iters))
    return {"ginstr_s": n_instr / (min(ms) * 1e-3) / 1e9, "ginstr_s_sustained": n_instr / (sum(ms) / len(ms) * 1e-3) / 1e9,
            "how": "csrc/peaks.cu: 8 independent DFMA chains per thread, 8 CTAs x 256 threads per SM, 20000 iterations, "
                   "best (burst) and mean (sustained) of 8 launches, CUDA events; one DFMA = one instruction"}


def timed_gk(torch, dist, world, st, K, W, proj, sampler=None):
    """W warm-up steps, then exactly K steps between barrier+synchronize, CUDA events on the launching stream.
    Returns (ms_total on this rank, [(launch ms) ...] in order A^T, A, A^T, A ..., launches)."""
    from trips_b200 import _lib
    from trips_b200 import kernels as KM

    spmv_events = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_events():
        # the single-call GK steps (tb200_gk_step_*_f64) record these around their two operator launches
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for e in evs:
            e.record()  # instantiate the CUDA event objects
        spmv_events.append((evs[0], evs[1]))
        spmv_events.append((evs[2], evs[3]))
        return evs

    def timed(fn):
        def run(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            spmv_events.append((e0, e1))
            return out
        return run

    for _ in range(W):
        st.step()
    barrier()
    orig_spmv = KM.spmv
    KM.spmv = timed(orig_spmv)
    KM.GK_STEP_EVENTS = step_events
    hooks = getattr(st, "bench_hooks", None)  # sharded state: wraps its two operator launches itself
    if hooks is not None:
        hooks(timed)
    if sampler is not None:
        sampler.mark_begin()
    launches0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        st.step()
    ev1.record()
    barrier()
    if sampler is not None:
        sampler.mark_end()
    KM.spmv = orig_spmv
    KM.GK_STEP_EVENTS = None
    if hooks is not None:
        hooks(None)
    ms = ev0.elapsed_time(ev1)
    return ms, [e0.elapsed_time(e1) for e0, e1 in spmv_events], _lib.launch_count - launches0


def stored_roofline(ms_AT, ms_A, nnz, m, n, val_bytes, hbm_peak, peak_src, traffic, kernel):
    bytes_A = (val_bytes + 4) * nnz + 8 * (m + 1) + 8 * n + 16 * m
    bytes_AT = (val_bytes + 4) * nnz + 8 * (n + 1) + 8 * m + 16 * n
    mean_ms = 0.5 * (ms_A + ms_AT)
    alg = 0.5 * (bytes_A + bytes_AT)
    ach = alg / (mean_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
            "peak_source": peak_src, "kernel": kernel, "launch_ms": mean_ms, "launch_ms_AT": ms_AT, "launch_ms_A": ms_A,
            "alg_bytes_per_launch": alg,
            "A_launch": {"alg_bytes": bytes_A, "achieved": bytes_A / (ms_A * 1e-3) / 1e9, "frac": bytes_A / (ms_A * 1e-3) / 1e9 / hbm_peak},
            "AT_launch": {"alg_bytes": bytes_AT, "achieved": bytes_AT / (ms_AT * 1e-3) / 1e9, "frac": bytes_AT / (ms_AT * 1e-3) / 1e9 / hbm_peak}}


def parity_block(torch, tb, O, A_full, nx, views, layout, dev):
    """Bit-for-bit evidence at the benchmarked geometry, outside the timed region.
    Eight angles spread over [0, pi) - axis aligned, steep, 45 degrees, shallow ascending / descending - of the FULL
    2048^2 image: the implicit operator restricted to them vs scipy on the oracle's NumPy statement of the matrix
    (forward, adjoint), three Golub-Kahan steps vs the oracle (correctly rounded norms), and the full 720-view operator's
    rows for those angles vs the restricted operator (the full launch takes other scheduling / alignment paths)."""
    import numpy as np

    t0 = time.perf_counter()
    n_det = O.ct_num_detectors(nx)
    sub = np.unique(np.array([0, views * 13 // 100, views // 4, views * 37 // 100, views // 2, views * 63 // 100,
                              3 * views // 4, views * 9 // 10]))
    A_s = O.ct_matrix(nx, O.ct_angles(views)[sub])
    op_s = tb.ParallelBeamCT(nx, views, angle_subset=sub, device=dev, layout=layout)
    m_s, n = A_s.shape
    rng = np.random.default_rng(2022)
    x = rng.standard_normal(n)
    u = rng.standard_normal(m_s)
    xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
    fwd = op_s.apply_dev(xd).cpu().numpy()
    adj = op_s.adjoint_dev(ud).cpu().numpy()
    fwd_ok = bool(np.array_equal(fwd, A_s @ x))
    adj_ok = bool(np.array_equal(adj, A_s.T @ u))
    b_s = A_s @ O.shepp_logan(nx).reshape(-1)
    steps = 3
    st = tb.GKState(op_s, torch.from_numpy(b_s).to(dev), steps)
    for _ in range(steps):
        st.step()
    _, al, be = st.scalars_host()
    with O.reductions("exact"):
        So = O.golub_kahan(A_s, b_s, steps)[1]
    al_o, be_o = np.diag(So), np.diag(So, -1)
    gk_ok = bool(np.array_equal(al, al_o) and np.array_equal(be, be_o))
    gk_dev = float(max(np.max(np.abs(al - al_o) / al_o), np.max(np.abs(be - be_o) / be_o)))
    # the full operator's rows of those angles
    full_rows_ok, adjointness = None, None
    if A_full is not None and A_full.shape[0] == views * n_det:
        yf = A_full.apply_dev(xd)
        rows = torch.from_numpy((sub[:, None] * n_det + np.arange(n_det)[None, :]).reshape(-1)).to(dev)
        full_rows_ok = bool(torch.equal(yf[rows], torch.from_numpy(fwd).to(dev)))
        uf = torch.from_numpy(rng.standard_normal(A_full.shape[0])).to(dev)
        zf = A_full.adjoint_dev(uf)
        lhs, rhs = float(torch.dot(yf, uf)), float(torch.dot(xd, zf))
        adjointness = abs(lhs - rhs) / max(abs(lhs), 1e-300)
    del op_s, st
    torch.cuda.empty_cache()
    return {"geometry": f"{nx}^2 image, angles {sub.tolist()} of {views} (m_s = {m_s}, nnz_s = {A_s.nnz})",
            "forward_bitwise_vs_scipy": fwd_ok, "adjoint_bitwise_vs_scipy": adj_ok,
            "gk_alpha_beta_bitwise_vs_oracle": gk_ok, "gk_steps": steps, "gk_alpha_beta_max_rel_dev": gk_dev,
            "full_operator_rows_bitwise_vs_subset": full_rows_ok, "full_operator_adjointness_rel": adjointness,
            "ok": bool(fwd_ok and adj_ok and gk_ok and full_rows_ok is not False
                       and (adjointness is None or adjointness < 1e-12)),
            "seconds": time.perf_counter() - t0}


def main():
    args = parse()
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    import trips_oracle as O

    nx, views = args.nx, args.views
    n_det = O.ct_num_detectors(nx)
    n = nx * nx
    m_full = views * n_det
    workload = f"cfg4 geometry: parallel-beam CT {nx}^2, {views} angles, {n_det} detectors, one Golub-Kahan step per step"
    # the part of `config` both arms share verbatim
    base_config = {"workload": workload, "nx": nx, "views": views, "n_det": n_det, "m": m_full, "n": n}
    hbm_peak, peak_src = peaks()

    if args.impl == "reference":
        return reference_arm(args, workload, base_config)

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: trips_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    import trips_b200 as tb
    from trips_b200 import _lib
    from trips_b200 import dist as tbdist

    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    layout = args.layout
    if layout == "auto":
        layout = "csr" if args.order != "sequential" else ("sell" if args.f32_storage else "implicit")
    if layout == "implicit" and (args.f32_storage or args.order != "sequential"):
        raise SystemExit("--layout implicit is the fp64 sequential-order path")
    _lib.check(_lib.lib().tb200_spmv_set_variant(args.variant))

    t_build = time.perf_counter()
    my_angles = tbdist.shard_angles(views, world, rank)
    A = tb.ParallelBeamCT(nx, views, angle_subset=my_angles if world > 1 else None, device=dev, layout=layout)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    if args.f32_storage:
        A = A.with_f32_storage()
        torch.cuda.empty_cache()
    if args.order != "sequential":
        A = A.with_order(args.order)
    m_loc, nnz_loc = A.shape[0], A.nnz
    nnz_t = torch.tensor([nnz_loc], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())

    # right-hand side: A x_true + 1 % noise, generated on the device per rank
    x_true = torch.from_numpy(O.shepp_logan(nx).reshape(-1)).to(dev)
    b = A.apply_dev(x_true)
    g = torch.Generator(device=dev)
    g.manual_seed(2022 + rank)
    noise = torch.randn(m_loc, dtype=torch.float64, device=dev, generator=g)
    b = b + 0.01 * float(b.norm()) / float(noise.norm()) * noise
    del noise, x_true

    K, W = args.steps, args.warmup
    exchange, exchange_fallback = None, None
    if world > 1:
        st = None
        if layout == "implicit" and args.exchange != "nccl":
            try:
                st = tbdist.BandShardedCT(nx, views, device=dev).gk_state(b, K + W)
            except Exception as exc:  # noqa: BLE001  (CUDA IPC unavailable in this container: use the NCCL transport)
                if args.exchange == "p2p":
                    raise
                exchange_fallback = f"{type(exc).__name__}: {exc}"[:300]
        if st is None:
            st = tbdist.DistGKState(A, b, K + W)
        exchange = st.exchange_name
    else:
        st = tb.GKState(A, b, K + W)
    proj = getattr(A, "projector", None)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, spmv_ms, launches = timed_gk(torch, dist, world, st, K, W, proj, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / K
    value = 1e3 / ms_per_step
    ms_AT = sum(spmv_ms[0::2]) / max(len(spmv_ms[0::2]), 1)
    ms_A = sum(spmv_ms[1::2]) / max(len(spmv_ms[1::2]), 1)
    share_A = ms_A * len(spmv_ms[1::2]) / ms
    share_AT = ms_AT * len(spmv_ms[0::2]) / ms
    beta0, al, be = st.scalars_host()
    assert np.isfinite(al).all() and np.isfinite(be).all() and (al > 0).all()

    # ------------------------------------------------------------------ roofline
    val_bytes = 4 if args.f32_storage else 8
    B_GK = 2 * (val_bytes + 4) * nnz + 8 * (m_full + 1) + 8 * (n + 1) + 48 * (m_full + n)
    gk_it = {"alg_bytes": B_GK, "achieved": B_GK / (ms_per_step * 1e-3) / 1e9 / world,
             "frac": B_GK / (ms_per_step * 1e-3) / 1e9 / world / hbm_peak,
             "note": "SURVEY 8(d) stored-matrix bytes per GK iteration / step time, per GPU; for the implicit layout a "
                     "streaming-equivalent speed, not a DRAM utilisation"}
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:  # noqa: BLE001
        traffic = {}
    single_cfg4 = world == 1 and (nx, views) == (2048, 720)
    fp64 = measure_fp64_peak(torch, dev)
    if layout == "implicit":
        # both launches are bound by the rate at which the SMs retire fp64 instructions, not by HBM: the honest
        # fraction is (fp64 instructions the launch executes) / time / (measured DFMA-chain rate).  The instruction
        # counts come from the projector (per candidate / per pixel-angle figures read off the SASS)
        # and are cross-checked against ncu's smsp__inst_executed_pipe_fp64 in profiles/.
        cnt = proj.fp64_instruction_counts()
        tr = traffic.get("cfg4_implicit_n1", {}) if single_cfg4 else {}
        bytes_A = 12 * nnz_loc + 8 * (m_loc + 1) + 8 * n + 16 * m_loc
        bytes_AT = 12 * nnz_loc + 8 * (n + 1) + 8 * m_loc + 16 * n

        def leg(name, kernel, ms_l, instr, alg_bytes, share, traffic_b):
            ach = instr / (ms_l * 1e-3) / 1e9
            return {"bound": "fp64-issue", "achieved": ach, "peak": fp64["ginstr_s"], "unit": "G fp64 instr/s",
                    "frac": ach / fp64["ginstr_s"], "traffic": traffic_b, "peak_source": fp64["how"],
                    "kernel": kernel, "launch_ms": ms_l, "fp64_instr_per_launch": instr, "share_of_step": share,
                    "hbm_equivalent": {"alg_bytes_per_launch": alg_bytes, "achieved": alg_bytes / (ms_l * 1e-3) / 1e9,
                                       "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (ms_l * 1e-3) / 1e9 / hbm_peak,
                                       "peak_source": peak_src,
                                       "note": "SURVEY 8(d) stored-matrix bytes / launch time: what a streaming SpMV would "
                                               "have to sustain to match; not a DRAM utilisation (see traffic)"}}

        legs = {"forward": leg("forward", cnt["forward_kernel"], ms_A, cnt["forward"], bytes_A, share_A, tr.get("A_launch_bytes")),
                "back_projection": leg("back", cnt["backproject_kernel"], ms_AT, cnt["backproject"], bytes_AT, share_AT,
                                       tr.get("AT_launch_bytes"))}
        dom = "forward" if ms_A >= ms_AT else "back_projection"
        other = "back_projection" if dom == "forward" else "forward"
        roofline = dict(legs[dom])
        roofline["other_launch"] = legs[other]
        roofline["fp64_peak"] = fp64
        roofline["gk_iteration"] = gk_it
        roofline["launch_ms_A"], roofline["launch_ms_AT"] = ms_A, ms_AT
    else:
        key = "cfg4_sell_sequential_n1" if (single_cfg4 and layout == "sell" and not args.f32_storage) else None
        kernel_name = {("sequential", "sell"): "spmv_sell_kernel (SELL-32-4, stored values)",
                       ("sequential", "csr"): "spmv_seq_tile_kernel", ("tree", "csr"): "spmv_warp_kernel"}[(args.order, layout)]
        roofline = stored_roofline(ms_AT, ms_A, nnz_loc, m_loc, n, val_bytes, hbm_peak, peak_src,
                                   traffic.get(key, {}).get("mean_bytes") if key else None, kernel_name)
        roofline["share_of_step"] = share_A + share_AT
        roofline["gk_iteration"] = gk_it
        roofline["fp64_peak"] = fp64

    # ------------------------------------------------------------------ parity (outside the timed region)
    parity = None
    if not args.no_parity and args.order == "sequential" and not args.f32_storage:
        if world == 1:
            if rank == 0:
                parity = parity_block(torch, tb, O, A, nx, views, layout, dev)
        else:
            # sharded run vs the same operator on ONE GPU: rank 0 rebuilds the whole problem and repeats the first steps
            parity = tbdist.sharded_parity_check(st, nx, views, layout, b, steps=min(10, K + W))

    # ------------------------------------------------------------------ e2e: reference-signature call, host buffers
    e2e = None
    if world > 1 and isinstance(st, tbdist.ShardedGKState):
        # every rank holds ITS rows of U and ITS band of V in pinned host memory; each step uploads u_k and v_{k-1},
        # runs the sharded step and downloads the new u, v and (alpha, beta)
        hu = torch.empty((args.e2e_steps + 3, st.m_loc), dtype=torch.float64).pin_memory()
        hv = torch.empty((args.e2e_steps + 3, st.n_band), dtype=torch.float64).pin_memory()
        hu[0].copy_(st.U.data[0])
        torch.cuda.synchronize()
        beta_prev = 0.0
        dt = 0.0
        for i in range(args.e2e_steps + 2):
            if i == 2:
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            al_i, be_i = st.host_step(hu[i], hv[i - 1] if i else None, beta_prev, hu[i + 1], hv[i])
            beta_prev = be_i
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item()) / args.e2e_steps
        e2e = {"value": 1.0 / dt, "unit": UNIT, "h2d_bytes_per_step": 8 * (st.m_loc + st.n_band),
               "d2h_bytes_per_step": 8 * (st.m_loc + st.n_band) + 16, "ms_per_step": dt * 1e3,
               "api": f"one golub_kahan_update per step with HOST (pinned) bases, sharded: per rank its rows of U and its band "
                      f"of V (ShardedGKState.host_step; bytes are per rank, x{world} ranks)"}
        st.close()
    del st
    torch.cuda.empty_cache()
    comm = tbdist.RowComm() if world > 1 else None
    if e2e is None:
        A_call = A if world == 1 else tbdist.ShardedRowsOperator(A, comm)
        b_host = b.cpu().numpy().reshape(-1, 1)
        bn = float(np.sqrt(comm.allreduce_(torch.tensor([float(b_host.T @ b_host)], dtype=torch.float64, device=dev)).item())) \
            if world > 1 else float(np.linalg.norm(b_host))
        U = b_host / bn
        S, V = np.empty(1), np.empty((n, 1))
        for _ in range(2):
            U, S, V = tb.golub_kahan_update(A_call, U, S, V, b200_comm=comm)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            U, S, V = tb.golub_kahan_update(A_call, U, S, V, b200_comm=comm)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        dt /= args.e2e_steps
        e2e = {"value": 1.0 / dt, "unit": UNIT, "h2d_bytes_per_step": 8 * (m_loc + n), "d2h_bytes_per_step": 8 * (m_loc + n) + 32,
               "ms_per_step": dt * 1e3,
               "api": "trips_b200.golub_kahan_update(A, U, S, V) with NumPy U, S, V (pinned host bases)"
                      + ("" if world == 1 else f"; per rank: its rows of U and the whole V (bytes are per rank, x{world} ranks)")}
        del U, V
        tb.release_host_buffers()

    # ------------------------------------------------------------------ secondary records (N = 1): the stored SpMV
    secondary = None
    if world == 1 and rank == 0 and not args.no_secondary and layout == "implicit":
        secondary = {}
        free_b = torch.cuda.mem_get_info(dev)[0]
        if 2 * 12.4 * nnz < free_b - 8e9:
            As = tb.ParallelBeamCT(nx, views, device=dev, layout="sell")
            for name, op, vb in (("stored_sell_f64", As, 8), ("stored_sell_f32_storage", None, 4)):
                if op is None:
                    op = As.with_f32_storage()
                    del As
                    torch.cuda.empty_cache()
                st2 = tb.GKState(op, b, K + W)
                ms2, sp2, _ = timed_gk(torch, None, 1, st2, K, W, None)
                aT, aA = sum(sp2[0::2]) / len(sp2[0::2]), sum(sp2[1::2]) / len(sp2[1::2])
                _, al2, be2 = st2.scalars_host()
                kk = min(len(al), len(al2))
                rec = {"metric": METRIC, "value": 1e3 / (ms2 / K), "unit": UNIT, "ms_per_step": ms2 / K,
                       "dtype": "f64" if vb == 8 else "f32-storage/f64-accumulate", "layout": "sell",
                       "roofline": stored_roofline(aT, aA, nnz, m_full, n, vb, hbm_peak, peak_src,
                                                   traffic.get("cfg4_sell_sequential_n1", {}).get("mean_bytes")
                                                   if (single_cfg4 and vb == 8) else None,
                                                   "spmv_sell_kernel (SELL-32-4, stored values: the north_star's CSR SpMV with "
                                                   "an explicitly stored transpose)"),
                       "gk_iteration": {"alg_bytes": 2 * (vb + 4) * nnz + 8 * (m_full + 1) + 8 * (n + 1) + 48 * (m_full + n)},
                       # same right-hand side as the primary record: the factors must agree bit for bit (fp64) or to the
                       # storage rounding (fp32 storage)
                       "alpha_beta_max_rel_dev_vs_primary": float(max(np.max(np.abs(al2[:kk] - al[:kk]) / al[:kk]),
                                                                      np.max(np.abs(be2[:kk] - be[:kk]) / be[:kk]))),
                       # (the recurrences run without reorthogonalisation: a storage rounding of 6e-8 per entry is
                       # amplified step by step exactly as any perturbation of the reference's own run is - DESIGN.md 2)
                       "alpha_beta_rel_dev_first_3_steps": float(max(np.max(np.abs(al2[:3] - al[:3]) / al[:3]),
                                                                     np.max(np.abs(be2[:3] - be[:3]) / be[:3])))}
                rec["gk_iteration"]["achieved"] = rec["gk_iteration"]["alg_bytes"] / (ms2 / K * 1e-3) / 1e9
                rec["gk_iteration"]["frac"] = rec["gk_iteration"]["achieved"] / hbm_peak
                secondary[name] = rec
                del st2
                torch.cuda.empty_cache()
            del op
            torch.cuda.empty_cache()
        else:
            secondary["skipped"] = f"stored pair needs {2 * 12.4 * nnz / 1e9:.0f} GB, {free_b / 1e9:.0f} GB free"

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import ref_loader

        sub = np.linspace(0, views, args.cpu_sample_views, endpoint=False).astype(int)
        op = tb.ParallelBeamCT(nx, views, angle_subset=sub, device=dev, layout="csr")
        A_s = op.to_scipy()  # device-built, bit-identical to the NumPy statement (tests); 30 s faster than building it on the host
        del op
        torch.cuda.empty_cache()
        b_s = A_s @ O.shepp_logan(nx).reshape(-1)
        r = cpu_gk(A_s, b_s, 6, 1, m_full, nnz, ref_loader.available())
        cpu = {"value": 1.0 / r["t_full_s"], "unit": UNIT, "cores": 1, "kind": r["kind"], "host_cpus": os.cpu_count(),
               "sample": f"{'reference' if r['kind'] == 'reference' else 'oracle'} golub_kahan_update (scipy csr/csc matvec, "
                         f"1 thread) on {len(sub)} of {views} angles of the full image (nnz {A_s.nnz:.3e}), "
                         f"{r['t_step_sample_s']:.3f} s per sample step; {r['formula']}",
               "extrapolation": r}

    if rank == 0:
        config = dict(base_config)
        config.update({"nnz": nnz, "parallelism": (f"u-space by angle, v-space by image band x{world}; exchange: {exchange}"
                                                   if world > 1 else "single GPU"),
                       "exchange_fallback": exchange_fallback, "spmv_order": args.order, "spmv_variant": args.variant, "layout": layout,
                       "matrix_bytes": (A.projector.nbytes if layout == "implicit" else 2 * (val_bytes + 4) * nnz),
                       "l2_note": (("matrix-free: per step the kernels read the image and the sinogram (L2 resident by "
                                    "design) and write them once; nothing else is streamed, so there is no cold input to "
                                    "flush" if A.projector.stored == 0 else
                                    "inputs (index stream per step) exceed L2 by >100x; no flush needed") if layout == "implicit"
                                   else "inputs (2 x 46 GB matrix streams per step) exceed L2 by >300x; no flush needed"),
                       "build_s": t_build})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64" if not args.f32_storage else "f32-storage/f64-accumulate", "data": "synthetic",
                "config": config, "roofline": roofline, "parity": parity, "cpu_baseline": cpu, "e2e": e2e,
                "secondary": secondary, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
