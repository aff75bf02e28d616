#!/usr/bin/env python
"""bench.py -- SO(3) reparameterize + Wigner-D action, fwd+bwd samples/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a kernels (one rank per GPU under torchrun)
    python bench.py --impl reference ...                     # the reference algorithm (oracle port) on the host CPU

Workload (BASELINE.json configs[4], the one the metric is quoted on): 2^24 samples in total, sharded
by batch over the ranks; per sample: fused reparameterize (+ wrapped log-density, k = 3) -> matrix ->
ZYZ Euler -> Wigner-D action (l <= 8, 10 channels) forward, then the full backward.  One *step* is one
pass over the rank's shard in micro-batches of 2^20 samples (the 54 GB output never needs to be
resident), followed by the only collective: an NCCL all-reduce of [loss, grad item_rep] (811 floats).

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same
pipeline through the public autograd API with mu/sigma in pinned HOST memory (the noise is generated
in the kernel, as the reference's module draws it itself), copies in the timed region; `roofline` =
the dominant kernel (Wigner backward) against the measured HBM copy peak, its duration taken from CUDA
events around every launch INSIDE the timed region (plain launches; `--graph` replays a CUDA graph
instead, same throughput) with the kernel-alone burst figure beside it; `cpu_baseline` = the oracle
port of the reference on this box's host cores, bounded sample; `ref_cuda_eager` = the same port as
eager PyTorch on cuda:0 (the reference's own execution model on this box); `parity` = a 4 096-sample
slice of the step against the FP64 oracle.

Inputs are generated per GLOBAL chunk of 2^18 samples (seed = f(chunk index)) and the upstream gradient
buffers are indexed by the global micro-batch number, so the loss and the item_rep gradient of a step are
the same numbers (to summation order) whatever the number of ranks.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "so3_reparam_wignerD_fwd_bwd_samples_per_sec"
UNIT = "samples/s"
TOTAL_SAMPLES = 1 << 24
MICRO = 1 << 20
L_MAX, CHANNELS, K_WIND = 8, 10, 3
FALLBACK_HBM_GBS = 6650.0
CPU_SAMPLE = 1 << 16        # bounded CPU sample (= BASELINE config 3's batch); a few seconds per step on 16 cores
GEN_CHUNK = 1 << 18         # inputs are generated per global chunk of this many samples: world-size independent
PARITY_SAMPLES = 4096
PHILOX_SEED = 20261018
TRAFFIC_PROFILE = os.path.join("profiles", "r02_wigner_bwd_traffic.json")     # ncu dram__bytes of the dominant kernel


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ reference arm / CPU baseline
def cpu_reference_run(samples, steps, warmup, threads, device="cpu"):
    """The reference's algorithm (oracle port, test infrastructure) for the same per-sample pipeline as eager PyTorch:
    on the host CPU (the reference arm / cpu_baseline) or, with ``device="cuda"``, on the GPU (ref_cuda_eager)."""
    from oracle import so3_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    mu = O.random_group_matrices(samples, generator=g).to(device)
    sigma = torch.nn.functional.softplus(torch.randn(samples, 3, generator=g)).to(device)
    eps = torch.randn(1, samples, 3, generator=g).to(device)
    item = torch.randn((L_MAX + 1) ** 2, CHANNELS, generator=g).to(device)
    gy = torch.randn(samples, (L_MAX + 1) ** 2 * CHANNELS, generator=g).to(device)
    glq = torch.randn(1, samples, generator=g).to(device)
    cuda = torch.device(device).type == "cuda"

    def step():
        m, s, it = mu.clone().requires_grad_(True), sigma.clone().requires_grad_(True), item.clone().requires_grad_(True)
        z, lq = O.so3_reparameterize(m, s, eps, K_WIND)
        ang = O.group_matrix_to_eazyz(z[0])
        y = O.action_net_forward(ang, it, L_MAX)
        loss = (y * gy).sum() + (lq * glq).sum()
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        if cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()                      # float(loss) inside synchronises: launch overhead is part of eager execution
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b) * 1e-3)
        else:
            t0 = time.perf_counter()
            step()
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = CPU_SAMPLE
    times = cpu_reference_run(sample, args.steps, max(1, min(args.warmup, 2)), threads)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.gpus), reference_arm_samples_per_step=sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d samples per step: a bounded sample of the 2^24-sample workload (the per-sample cost of the reference's "
                                   "algorithm does not depend on the batch); oracle port of the reference (the reference is Python and does not "
                                   "travel to the GPU box; the port is pinned to it by tests/golden), torch CPU fp32, %d threads" % (sample, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[4]: 2^24 samples, fused SO3 reparameterize (k=3) -> ZYZ Euler -> Wigner-D action "
                        "(l<=8, 81-dim, 10 channels), fwd+bwd, batch-sharded",
            "total_samples": TOTAL_SAMPLES, "micro_batch": MICRO, "degrees": L_MAX, "rep_copies": CHANNELS, "k": K_WIND,
            "parallelism": "dp%d" % n_gpus, "collective": "all-reduce of [loss, grad item_rep] (811 f32) per step",
            "l2": "inputs_exceed_l2 (every micro-batch streams > 6.8 GB; no reuse between timed iterations)"}


# ------------------------------------------------------------------------------ our arm
def gen_inputs(lo, hi, dev, lt):
    """mu, sigma, eps, g_log_q of the global samples [lo, hi): generated per GEN_CHUNK-aligned chunk from a generator seeded
    with the chunk index (uniform rotations as lie_tools.py:256-267, sigma = softplus(randn) as reparameterize.py:121), so
    a sample's inputs do not depend on how many ranks share the work."""
    assert lo % GEN_CHUNK == 0 and (hi - lo) % GEN_CHUNK == 0, "shards are multiples of the generation chunk"
    n = hi - lo
    mu = torch.empty(n, 3, 3, device=dev)
    sigma, eps, glq = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev), torch.empty(n, device=dev)
    for c in range(lo // GEN_CHUNK, hi // GEN_CHUNK):
        g = torch.Generator(device=dev).manual_seed(0x5EED0000 + c)
        o = c * GEN_CHUNK - lo
        u1, u2, u3 = torch.rand(3, GEN_CHUNK, device=dev, generator=g)
        q = torch.stack([torch.sqrt(1 - u1) * torch.sin(2 * math.pi * u2), torch.sqrt(1 - u1) * torch.cos(2 * math.pi * u2),
                         torch.sqrt(u1) * torch.sin(2 * math.pi * u3), torch.sqrt(u1) * torch.cos(2 * math.pi * u3)], 1)
        mu[o:o + GEN_CHUNK] = lt.quaternions_to_group_matrix(q)
        sigma[o:o + GEN_CHUNK] = torch.nn.functional.softplus(torch.randn(GEN_CHUNK, 3, device=dev, generator=g))
        eps[o:o + GEN_CHUNK] = torch.randn(GEN_CHUNK, 3, device=dev, generator=g)
        glq[o:o + GEN_CHUNK] = torch.randn(GEN_CHUNK, device=dev, generator=g)
    return mu, sigma, eps, glq


def step_loss(item, g_item, log_q, g_log_q):
    a = (item.double() * g_item.double()).sum()
    b = (log_q.view(-1, GEN_CHUNK) * g_log_q.view(-1, GEN_CHUNK)).sum(1).double().sum()
    return (a + b).float()


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity) BEFORE the pinned staging buffers are allocated:
    first-touch then places them on the GPU's NUMA node and eight ranks do not pull their inputs across the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return len(os.sched_getaffinity(0))
    except Exception:                                         # noqa: BLE001 -- an optimisation, never a requirement
        return None


def parity_block(step_cls, lt, mu, sigma, eps, glq, item, gy0, dev):
    """A PARITY_SAMPLES slice of the step (explicit-eps kernels: angles, log_q, y, g_mu, g_sigma, g_item) against the FP64
    oracle on the CPU, outside the timed region; the oracle run in FP32 (= the reference's own arithmetic) beside it."""
    from oracle import so3_oracle as O
    n, M = PARITY_SAMPLES, (L_MAX + 1) ** 2
    st = step_cls(n, n, L_MAX, CHANNELS, K_WIND, device=dev)
    m, s, e, gl, gy = (t[:n].contiguous() for t in (mu, sigma, eps, glq, gy0))
    lq, y = torch.empty(n, device=dev), torch.empty(n, M * CHANNELS, device=dev)
    gm, gs = torch.empty(n, 3, 3, device=dev), torch.empty(n, 3, device=dev)
    st.latent_forward(m, s, e, lq)
    st.decode_forward(0, n, item, y)
    st.decode_backward(0, n, item, gy)
    st.latent_backward(m, s, e, gl, gm, gs)
    torch.cuda.synchronize()
    ours = {"angles": st.angles, "log_q": lq, "y": y, "g_mu": gm, "g_sigma": gs, "g_item_rep": st.g_item}
    ours = {k: v.detach().double().cpu() for k, v in ours.items()}

    def oracle(dtype):
        mm, ss, it = (t.detach().cpu().to(dtype).requires_grad_(True) for t in (m, s, item))
        z, q = O.so3_reparameterize(mm, ss, e.cpu().to(dtype).view(1, n, 3), K_WIND)
        ang = O.group_matrix_to_eazyz(z[0])
        yy = O.action_net_forward(ang, it, L_MAX)
        ((yy * gy.cpu().to(dtype)).sum() + (q[0] * gl.cpu().to(dtype)).sum()).backward()
        out = {"angles": ang, "log_q": q[0], "y": yy, "g_mu": mm.grad, "g_sigma": ss.grad, "g_item_rep": it.grad}
        return {k: v.detach().double() for k, v in out.items()}
    ref64, ref32 = oracle(torch.float64), oracle(torch.float32)
    rtol = atol0 = 1e-5
    per, tot_out, tot_ref_out, tot_n, worst_abs, worst_rel = {}, 0, 0, 0, 0.0, 0.0
    for k in ours:
        a, b, c = ours[k].reshape(-1), ref64[k].reshape(-1), ref32[k].reshape(-1)
        ok = torch.isfinite(b)
        a, b, c = a[ok], b[ok], c[ok]
        atol = atol0 * max(1.0, float(b.pow(2).mean().sqrt()))            # the tests' rule: atol scaled by the tensor's rms
        err, err32 = (a - b).abs(), (c - b).abs()
        out, out32 = int((err > atol + rtol * b.abs()).sum()), int((err32 > atol + rtol * b.abs()).sum())
        rel = float((err / b.abs().clamp_min(1.0)).max())
        per[k] = {"max_abs": float(err.max()), "max_rel": rel, "frac_outside_tol": out / a.numel(),
                  "reference_fp32_frac_outside_tol": out32 / a.numel(), "reference_fp32_max_abs": float(err32.max())}
        tot_out, tot_ref_out, tot_n = tot_out + out, tot_ref_out + out32, tot_n + a.numel()
        worst_abs, worst_rel = max(worst_abs, float(err.max())), max(worst_rel, rel)
    return {"samples": n, "against": "float64 oracle (CPU), same inputs and noise", "tol": "|a-b| <= 1e-5*max(1,rms) + 1e-5*|b|; max_rel = max |a-b| / max(|b|, 1)",
            "max_abs": worst_abs, "max_rel": worst_rel, "frac_outside_tol": tot_out / tot_n,
            "reference_fp32_frac_outside_tol": tot_ref_out / tot_n, "per_output": per}


def run_ours(args):
    import torch.distributed as dist
    from lie_vae_b200 import _build
    from lie_vae_b200 import dist as lvdist
    from lie_vae_b200.pipeline import FusedSO3ActionStep, KERNELS, algorithmic_bytes
    import lie_vae_b200.lie_tools as lt
    import lie_vae_b200.reparameterize as rp
    from lie_vae_b200 import _ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this arm has no CPU fallback (use --impl reference for the CPU baseline)")
    affinity0 = os.sched_getaffinity(0)
    cpus_bound = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if _build.is_stale():
        if rank == 0:
            _build.build()
        if world > 1:
            dist.barrier()

    total = args.samples
    n_loc = total // world
    micro = min(args.micro, n_loc)
    n_micro = n_loc // micro
    n_loc = n_micro * micro
    if n_loc % GEN_CHUNK or micro % GEN_CHUNK:
        raise SystemExit("bench.py: shard (%d) and micro-batch (%d) must be multiples of %d samples" % (n_loc, micro, GEN_CHUNK))
    lo, hi = lvdist.shard_bounds(n_loc * world, world, rank, micro)        # contiguous slice of the global batch
    assert hi - lo == n_loc
    M = (L_MAX + 1) ** 2

    mu, sigma, eps, glq = gen_inputs(lo, hi, dev, lt)
    gs = torch.Generator(device=dev).manual_seed(0xA11CE)
    item = torch.randn(M, CHANNELS, device=dev, generator=gs)                 # the same item_rep on every rank
    NBUF = 3
    # upstream gradient of y (stand-in for the decoder): buffer j serves the GLOBAL micro-batches j, j + NBUF, ...
    gy = [torch.randn(micro, M * CHANNELS, device=dev, generator=torch.Generator(device=dev).manual_seed(0xD0 + j)) for j in range(NBUF)]
    g_first = (lo // micro)                                                   # global index of this rank's first micro-batch
    y = [torch.empty(micro, M * CHANNELS, device=dev) for _ in range(2)]
    log_q = torch.empty(n_loc, device=dev)
    g_mu = torch.empty(n_loc, 3, 3, device=dev)
    g_sigma = torch.empty(n_loc, 3, device=dev)
    red = torch.zeros(1 + M * CHANNELS, device=dev)
    step_obj = FusedSO3ActionStep(n_loc, micro, L_MAX, CHANNELS, K_WIND, device=dev)
    g_item = step_obj.g_item          # the micro-batches accumulate into one gradient (no per-micro-batch add kernel)

    def local_step():
        g_item.zero_()
        step_obj.latent_forward(mu, sigma, eps, log_q)
        for i in range(n_micro):
            a, b = i * micro, (i + 1) * micro
            step_obj.decode_forward(a, b, item, y[i % 2])
            step_obj.decode_backward(a, b, item, gy[(g_first + i) % NBUF], accumulate=True)
        step_obj.latent_backward(mu, sigma, eps, glq, g_mu, g_sigma)
        # loss = sum(y * g_y) + sum(log_q * g_lq);  y is linear in item_rep, so sum(y * g_y) = <item_rep, grad item_rep>.
        # The two terms are large and cancel: partial sums per GLOBAL generation chunk, then float64, so that the scalar does
        # not depend on how the batch is cut into ranks beyond the rounding of grad item_rep itself.
        lvdist.pack_reduction(step_loss(item, g_item, log_q, glq), g_item, out=red)

    # The rank-local part of a step is a fixed sequence of launches on caller-owned buffers.  Default: plain launches with a
    # CUDA event on either side of each of the four kernels, so that the per-kernel durations of the roofline block are
    # measured INSIDE the timed region (sustained clocks; the host runs far ahead of the GPU at these kernel sizes).  --graph
    # captures the sequence once and replays it (same throughput, measured; the per-kernel durations then come from a second,
    # un-timed pass of plain launches).  The one collective stays outside either way.
    graph = None
    if args.graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                local_step()                                  # lazy initialisations (cuBLAS handle, smem opt-in) before capture
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                local_step()
        except Exception as e:                                # noqa: BLE001 -- fall back to plain launches, say so in the line
            sys.stderr.write("bench.py: CUDA graph capture failed (%s); using plain launches\n" % e)
            graph = None
            torch.cuda.synchronize()

    def all_reduce_step_result():
        """The path's only collective (lie_vae_b200.dist): sum of [loss, grad item_rep] over the ranks."""
        return lvdist.unpack_reduction(lvdist.allreduce_packed(red), (M, CHANNELS))

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            local_step()
        all_reduce_step_result()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler is a subprocess: start it BEFORE the warm-up steps, so that spawning it does not leave the GPU idle (and
    # its clocks ramping back up) between the warm-up and the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        one_step()
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps - 1)]     # between the steps: where the time goes
    live_events = graph is None
    if live_events:
        step_obj.enable_kernel_timing(True)
    t0.record()
    for i in range(args.steps):
        one_step()
        if i < args.steps - 1:
            marks[i].record()
    t1.record()
    barrier()
    elapsed_ms = t0.elapsed_time(t1)
    edges = [t0] + marks + [t1]
    per_step_ms = [round(a.elapsed_time(b), 3) for a, b in zip(edges[:-1], edges[1:])]
    loss_value = float(red[0])
    g_item_norm = float(red[1:].double().norm())
    loss_term_a = float((item.double().view(-1) * red[1:].double()).sum())
    if live_events:
        torch.cuda.synchronize()
        kern_ms = {k: [a.elapsed_time(b) for a, b in v] for k, v in step_obj.events.items()}
        plain_step_ms = None
    else:
        # graph replay: per-kernel durations from a second pass of as many steps, plain launches bracketed by CUDA events
        step_obj.enable_kernel_timing(True)
        plain = []
        for _ in range(max(2, args.steps)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            local_step()
            b.record()
            plain.append((a, b))
        torch.cuda.synchronize()
        kern_ms = {k: [a.elapsed_time(b) for a, b in v] for k, v in step_obj.events.items()}
        plain_step_ms = sum(a.elapsed_time(b) for a, b in plain) / len(plain)
    step_obj.enable_kernel_timing(False)
    # the dominant kernel timed alone (burst: a short idle, then 8 back-to-back launches): what it does outside a power-capped
    # sustained step; reported next to the in-step figure, never instead of it
    torch.cuda.synchronize()
    time.sleep(0.25)
    ba, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_burst = 8
    step_obj.decode_backward(0, micro, item, gy[0], accumulate=True)
    ba.record()
    for j in range(n_burst):
        step_obj.decode_backward(0, micro, item, gy[j % NBUF], accumulate=True)
    bb.record()
    torch.cuda.synchronize()
    alone_ms = ba.elapsed_time(bb) / n_burst

    # ---- end to end through the public autograd API, inputs in pinned host memory ------------------
    # What a caller of the module hands over per sample is mu (36 B) and sigma (12 B); the noise is drawn inside the kernel
    # (Philox counter = global sample index: world-size independent), as the reference's module draws it itself.  Micro-batches
    # of n_loc / 8 (at most 2^20) samples through a ring of four staging buffers keep several copies in flight per rank.
    split = max(1, args.e2e_split)
    e_micro = max(GEN_CHUNK, min(args.e2e_micro, n_loc // split if n_loc >= split * GEN_CHUNK else n_loc))
    while n_loc % e_micro or micro % e_micro and e_micro % micro:
        e_micro //= 2
    e_n_micro = n_loc // e_micro
    NST = 4

    def gy_for(i):
        """upstream gradient rows of the rank's e2e micro-batch i = the rows the resident pass pairs with these samples"""
        a = i * e_micro
        buf = gy[(g_first + a // micro) % NBUF]
        off = a % micro
        return buf[off:off + e_micro] if e_micro <= micro else torch.cat([gy[(g_first + a // micro + r) % NBUF] for r in range(e_micro // micro)])
    mu_h, sg_h = (t.cpu().pin_memory() for t in (mu, sigma))
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [[torch.empty(e_micro, 3, 3, device=dev), torch.empty(e_micro, 3, device=dev)] for _ in range(NST)]
    ready = [torch.cuda.Event() for _ in range(NST)]
    freed = [torch.cuda.Event() for _ in range(NST)]
    out_h = torch.empty(1 + M * CHANNELS).pin_memory()
    item_p = item.clone().requires_grad_(True)
    lq_terms = torch.empty(n_loc // GEN_CHUNK, device=dev)          # per generation chunk: sum(log_q * g_log_q), summed in float64 at the end
    lq_prod = torch.empty(e_micro // GEN_CHUNK, GEN_CHUNK, device=dev)
    cpg = e_micro // GEN_CHUNK
    host_s = [0.0]

    def issue(i):
        b = i % NST
        sl = slice(i * e_micro, (i + 1) * e_micro)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            stage[b][0].copy_(mu_h[sl], non_blocking=True)
            stage[b][1].copy_(sg_h[sl], non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_step(prefetched, prefetch_next):
        """One step; every micro-batch's mu / sigma travel host -> device inside it.  ``prefetch_next``: the copies of the NEXT
        step's first micro-batches are issued before this step's result is read back (a loader that stays one batch ahead),
        so only the first timed step pays the fill of the staging ring."""
        t_host = time.perf_counter()
        item_p.grad = None
        main = torch.cuda.current_stream()
        if not prefetched:
            for b in range(NST):
                freed[b].record(main)
            for i in range(min(NST - 1, e_n_micro)):
                issue(i)
        for i in range(e_n_micro):
            b = i % NST
            if i + NST - 1 < e_n_micro:
                issue(i + NST - 1)
            main.wait_event(ready[b])
            m = stage[b][0].requires_grad_(True)
            s = stage[b][1].requires_grad_(True)
            ang3, lq = rp.so3_reparameterize_philox(m, s, 1, K_WIND, PHILOX_SEED, lo + i * e_micro, euler=True)
            yy = _ops.wigner_apply(ang3[0], item_p, 0, L_MAX, False)
            # the decoder that would consume y is outside the hot path: its gradient g_y (and g_log_q) is handed
            # to autograd directly, exactly as a downstream module's backward would
            glq_i = glq[i * e_micro:(i + 1) * e_micro]
            torch.autograd.backward([yy, lq], [gy_for(i).view(e_micro, M, CHANNELS), glq_i.view(1, e_micro)])
            torch.mul(lq.detach().view(cpg, GEN_CHUNK), glq_i.view(cpg, GEN_CHUNK), out=lq_prod)
            torch.sum(lq_prod, 1, out=lq_terms[i * cpg:(i + 1) * cpg])
            stage[b][0].grad = None
            stage[b][1].grad = None
            stage[b][0].requires_grad_(False)
            stage[b][1].requires_grad_(False)
            freed[b].record(main)
        if prefetch_next:
            for i in range(min(NST - 1, e_n_micro)):
                issue(i)
        # loss = sum(y * g_y) + sum(log_q * g_lq), with sum(y * g_y) = <item_rep, grad item_rep> (y is linear in item_rep)
        lvdist.pack_reduction((lq_terms.double().sum() + (item_p.detach().double() * item_p.grad.double()).sum()).float(), item_p.grad, out=red)
        all_reduce_step_result()
        out_h.copy_(red, non_blocking=True)
        host_s[0] += time.perf_counter() - t_host                  # host time to issue the step (before waiting for the GPU)
        torch.cuda.current_stream().synchronize()
        return float(out_h[0])

    e2e_steps = max(1, args.e2e_steps)
    e2e_step(False, False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host_s[0] = 0.0
    for it in range(e2e_steps):
        e2e_loss = e2e_step(it > 0, it + 1 < e2e_steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks -----------------------------------------------------------------------------
    tt = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(tt[0]), float(tt[1])
    samples_per_step = n_loc * world
    ms_per_step = elapsed_ms / args.steps
    value = samples_per_step / (ms_per_step / 1e3)

    if rank == 0:
        peak, peak_src = measured_peaks()
        abytes = algorithmic_bytes(L_MAX, CHANNELS)
        kernels = {}
        for k in KERNELS:
            avg_ms = sum(kern_ms[k]) / len(kern_ms[k])
            per_launch = micro if k.startswith("wigner") else n_loc
            gbs = abytes[k] * per_launch / (avg_ms * 1e-3) / 1e9
            kernels[k] = {"avg_ms": round(avg_ms, 4), "samples_per_launch": per_launch, "bytes_per_sample": abytes[k],
                          "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        dom = "wigner_bwd"
        # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel: parsed from the committed ncu summary of this
        # round's kernel (per launch there; scaled to this run's samples per launch), null if the file is missing
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, TRAFFIC_PROFILE)) as f:
                tp = json.load(f)
            traffic = (tp["dram_bytes_read"] + tp["dram_bytes_write"]) * (micro / float(tp["samples_per_launch"]))
            traffic_src = "%s (%s, ncu --set full, not measured in this run)" % (TRAFFIC_PROFILE, tp.get("kernel", "?"))
        except Exception:                                      # noqa: BLE001
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), total_samples=samples_per_step, micro_batch=micro),
            "clocks": clocks,
            "e2e": {"value": samples_per_step / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "h2d_bytes_per_step": n_loc * 48, "d2h_bytes_per_step": 4 * (1 + M * CHANNELS),
                    "micro_batch": e_micro, "copies_in_flight": NST - 1, "cpus_bound_to_gpu_numa_node": cpus_bound,
                    "host_issue_ms_per_step": round(host_s[0] / e2e_steps * 1e3, 3),
                    "api": "so3_reparameterize_philox(euler) -> WignerApply (torch.autograd); pinned host mu/sigma (48 B/sample), noise "
                           "generated in the kernel; 4-deep staging ring, the next step's first copies issued before the result is read back"},
            "gpu_launches": (step_obj.LAUNCHES_PER_MICROBATCH * n_micro + step_obj.LAUNCHES_PER_SHARD) * args.steps,
            "launch_mode": "cuda_graph_replay" if graph is not None else "plain",
            "roofline": {"bound": "hbm", "kernel": "wigner_bwd_dg_kernel<Cfg8BP> (+ wigner_reduce_partials)", "achieved": kernels[dom]["gbs"], "peak": peak,
                         "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes[dom] * micro,
                         "timed": "CUDA events around every launch inside the timed region" if graph is None else
                                  "second pass of plain launches after the graph-replayed timed region",
                         # the same kernel alone (0.25 s idle, then 8 back-to-back launches): not power-capped
                         "kernel_alone": {"avg_ms": round(alone_ms, 4), "gbs": round(abytes[dom] * micro / (alone_ms * 1e-3) / 1e9, 1),
                                          "frac": round(abytes[dom] * micro / (alone_ms * 1e-3) / 1e9 / peak, 4)}},
            "pipeline_roofline": {"bytes_per_sample": 6656, "achieved_gbs": round(value / world * 6656 / 1e9, 1),
                                  "frac": round(value / world * 6656 / 1e9 / peak, 4)},
            "kernels": kernels,
            # where a step's time goes: the four kernels' launches, summed (the rest is the reduction of the partial item_rep
            # gradients, the loss terms and the packing of the all-reduce buffer); the timed step
            "step_breakdown_ms": {"sum_of_kernel_launches": round(sum(sum(v) / len(v) * (n_micro if k.startswith("wigner") else 1) for k, v in kern_ms.items()), 3),
                                  "local_step_plain_launches_second_pass": None if plain_step_ms is None else round(plain_step_ms, 3),
                                  "timed_step": round(ms_per_step, 3)},
            "per_step_ms": per_step_ms if len(per_step_ms) <= 64 else per_step_ms[:64],
            "loss": loss_value, "g_item_rep_norm": g_item_norm, "e2e_loss": e2e_loss,
            # the loss is the sum of two large cancelling terms: compare runs relative to the terms, not to their difference
            "loss_terms": {"item_rep_dot_grad": loss_term_a, "log_q_dot_g_log_q": loss_value - loss_term_a},
        }
        if not args.no_parity:
            line["parity"] = parity_block(FusedSO3ActionStep, lt, mu, sigma, eps, glq, item, gy[g_first % NBUF], dev)
        if not args.no_cpu_baseline and world == 1:
            os.sched_setaffinity(0, affinity0)          # the CPU baseline gets every host core again
            threads = os.cpu_count() or 1
            sample = CPU_SAMPLE
            times = cpu_reference_run(sample, 3, 1, threads)
            best = min(times)
            line["cpu_baseline"] = {"value": sample / best, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d samples (bounded sample of the workload), best of 3 after 1 warm-up, torch CPU fp32; "
                                              "oracle port: the Python reference does not travel to the GPU box" % sample}
            # the same port as eager PyTorch on this GPU: the reference's own execution model on the same box
            g_times = cpu_reference_run(sample, 3, 1, threads, device=str(dev))
            line["ref_cuda_eager"] = {"value": sample / min(g_times), "unit": UNIT, "sample": "%d samples, best of 3 after 1 warm-up, "
                                      "oracle port in eager PyTorch fp32 on %s (CUDA events)" % (sample, dev)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=TOTAL_SAMPLES, help="global samples per step")
    ap.add_argument("--micro", type=int, default=MICRO)
    ap.add_argument("--e2e-micro", type=int, default=1 << 20, help="largest micro-batch of the autograd-API end-to-end pass")
    ap.add_argument("--e2e-steps", type=int, default=10, help="timed steps of the end-to-end pass")
    ap.add_argument("--e2e-split", type=int, default=8, help="micro-batches per rank and step of the end-to-end pass (if the shard allows)")
    ap.add_argument("--no-parity", action="store_true", help="skip the 4096-sample parity block against the float64 oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay a CUDA graph of the rank-local step instead of plain launches")
    ap.add_argument("--no-graph", action="store_true", help="(default since round 2; accepted for old command lines)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: anything a library writes to file descriptor 1 (NCCL prints its
    # version banner there) is sent to stderr, and the result goes to the saved descriptor
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
