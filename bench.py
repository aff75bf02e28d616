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
pipeline through the public autograd API with mu/sigma/eps in pinned HOST memory, copies in the
timed region; `roofline` = the dominant kernel (Wigner backward) against the measured HBM copy peak;
`cpu_baseline` = the oracle port of the reference on this box's host cores, bounded sample.
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
def cpu_reference_run(samples, steps, warmup, threads):
    """The reference's algorithm (oracle port, test infrastructure) for the same per-sample pipeline on the CPU."""
    from oracle import so3_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    mu = O.random_group_matrices(samples, generator=g)
    sigma = torch.nn.functional.softplus(torch.randn(samples, 3, generator=g))
    eps = torch.randn(1, samples, 3, generator=g)
    item = torch.randn((L_MAX + 1) ** 2, CHANNELS, generator=g)
    gy = torch.randn(samples, (L_MAX + 1) ** 2 * CHANNELS, generator=g)
    glq = torch.randn(1, samples, generator=g)

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
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d samples per step (bounded sample of the 2^24 workload), torch CPU fp32, %d threads" % (sample, threads)},
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
def run_ours(args):
    import torch.distributed as dist
    from lie_vae_b200 import _build
    from lie_vae_b200.pipeline import FusedSO3ActionStep, KERNELS, algorithmic_bytes
    import lie_vae_b200.lie_tools as lt
    import lie_vae_b200.reparameterize as rp
    from lie_vae_b200 import _ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this arm has no CPU fallback (use --impl reference for the CPU baseline)")
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
    micro = min(args.micro, total // world)
    n_loc = total // world
    n_micro = n_loc // micro
    n_loc = n_micro * micro
    M = (L_MAX + 1) ** 2

    torch.manual_seed(1234 + rank)
    mu = lt.random_group_matrices(n_loc, device=dev)
    sigma = torch.nn.functional.softplus(torch.randn(n_loc, 3, device=dev))
    eps = torch.randn(n_loc, 3, device=dev)
    glq = torch.randn(n_loc, device=dev)
    item = torch.randn(M, CHANNELS, device=dev)
    NBUF = 3
    gy = [torch.randn(micro, M * CHANNELS, device=dev) for _ in range(NBUF)]       # upstream gradient of y (stand-in for the decoder)
    y = [torch.empty(micro, M * CHANNELS, device=dev) for _ in range(2)]
    log_q = torch.empty(n_loc, device=dev)
    g_mu = torch.empty(n_loc, 3, 3, device=dev)
    g_sigma = torch.empty(n_loc, 3, device=dev)
    g_item = torch.zeros(M, CHANNELS, device=dev)
    red = torch.zeros(1 + M * CHANNELS, device=dev)
    step_obj = FusedSO3ActionStep(n_loc, micro, L_MAX, CHANNELS, K_WIND, device=dev)

    g_item = step_obj.g_item          # the micro-batches accumulate into one gradient (no per-micro-batch add kernel)

    def local_step():
        g_item.zero_()
        step_obj.latent_forward(mu, sigma, eps, log_q)
        for i in range(n_micro):
            lo, hi = i * micro, (i + 1) * micro
            step_obj.decode_forward(lo, hi, item, y[i % 2])
            step_obj.decode_backward(lo, hi, item, gy[i % NBUF], accumulate=True)
        step_obj.latent_backward(mu, sigma, eps, glq, g_mu, g_sigma)
        # loss = sum(y * g_y) + sum(log_q * g_lq);  y is linear in item_rep, so sum(y * g_y) = <item_rep, grad item_rep>
        red[0] = (item * g_item).sum() + torch.dot(log_q, glq)
        red[1:] = g_item.view(-1)

    # The rank-local part of a step is a fixed sequence of ~200 launches on caller-owned buffers: capture it once in a CUDA
    # graph and replay it (no per-launch CPU cost, no gaps between the kernels); the one collective stays outside.
    graph = None
    if not args.no_graph:
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

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            local_step()
        if world > 1:
            dist.all_reduce(red)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        one_step()
    t1.record()
    barrier()
    elapsed_ms = t0.elapsed_time(t1)
    loss_value = float(red[0])
    # per-kernel durations: a separate, un-timed pass of plain launches bracketed by CUDA events on the launching stream
    step_obj.enable_kernel_timing(True)
    for _ in range(2):
        local_step()
    torch.cuda.synchronize()
    kern_ms = {k: [a.elapsed_time(b) for a, b in v] for k, v in step_obj.events.items()}
    step_obj.enable_kernel_timing(False)

    # ---- end to end through the public autograd API, inputs in pinned host memory ------------------
    # larger micro-batches than the resident pass: a torch.autograd round trip costs ~0.15 ms of CPU time per Function pair,
    # so the tape-driven path amortises it over 2^20 samples (y / g_y of 3.4 GB each are still far from resident for 2^24)
    e_micro = max(micro, min(args.e2e_micro, n_loc))
    while n_loc % e_micro:
        e_micro //= 2
    e_n_micro = n_loc // e_micro
    ratio = e_micro // micro       # the same upstream gradients as the resident pass, laid out for the larger micro-batches
    gy_e = gy if ratio == 1 else [torch.cat([gy[(ratio * j + r) % NBUF] for r in range(ratio)]) for j in range(NBUF)]
    mu_h, sg_h, ep_h = (t.cpu().pin_memory() for t in (mu, sigma, eps.view(1, n_loc, 3)))
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [[torch.empty(e_micro, 3, 3, device=dev), torch.empty(e_micro, 3, device=dev), torch.empty(1, e_micro, 3, device=dev)]
             for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    out_h = torch.empty(1 + M * CHANNELS).pin_memory()
    item_p = item.clone().requires_grad_(True)

    def e2e_step():
        item_p.grad = None
        loss_acc = torch.zeros((), device=dev)
        main = torch.cuda.current_stream()

        def issue(i):
            b = i % 2
            sl = slice(i * e_micro, (i + 1) * e_micro)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[b])
                stage[b][0].copy_(mu_h[sl], non_blocking=True)
                stage[b][1].copy_(sg_h[sl], non_blocking=True)
                stage[b][2].copy_(ep_h[:, sl], non_blocking=True)
                ready[b].record(copy_stream)
        for b in range(2):
            freed[b].record(main)
        issue(0)
        for i in range(e_n_micro):
            b = i % 2
            if i + 1 < e_n_micro:
                issue(i + 1)
            main.wait_event(ready[b])
            m = stage[b][0].requires_grad_(True)
            s = stage[b][1].requires_grad_(True)
            ang3, lq = rp.so3_reparameterize_eazyz(m, s, stage[b][2], K_WIND)
            ang = ang3[0]
            yy = _ops.WignerApply.apply(ang, item_p, 0, L_MAX, False)
            # the decoder that would consume y is outside the hot path: its gradient g_y (and g_log_q) is handed
            # to autograd directly, exactly as a downstream module's backward would
            glq_i = glq[i * e_micro:(i + 1) * e_micro]
            torch.autograd.backward([yy, lq], [gy_e[i % len(gy_e)].view(e_micro, M, CHANNELS), glq_i.view(1, e_micro)])
            loss_acc += torch.dot(lq.detach()[0], glq_i)
            stage[b][0].grad = None
            stage[b][1].grad = None
            stage[b][0].requires_grad_(False)
            stage[b][1].requires_grad_(False)
            freed[b].record(main)
        # loss = sum(y * g_y) + sum(log_q * g_lq), with sum(y * g_y) = <item_rep, grad item_rep> (y is linear in item_rep)
        red[0] = loss_acc + (item_p.detach() * item_p.grad).sum()
        red[1:] = item_p.grad.view(-1)
        if world > 1:
            dist.all_reduce(red)
        out_h.copy_(red, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(out_h[0])

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_loss = e2e_step()
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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), total_samples=samples_per_step, micro_batch=micro),
            "clocks": clocks,
            "e2e": {"value": samples_per_step / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": n_loc * 60, "d2h_bytes_per_step": 4 * (1 + M * CHANNELS),
                    "micro_batch": e_micro,
                    "api": "so3_reparameterize_eazyz -> WignerApply (torch.autograd), pinned host mu/sigma/eps, double-buffered copies"},
            "gpu_launches": (step_obj.LAUNCHES_PER_MICROBATCH * n_micro + step_obj.LAUNCHES_PER_SHARD) * args.steps,
            "launch_mode": "cuda_graph_replay" if graph is not None else "plain",
            "roofline": {"bound": "hbm", "kernel": "wigner_bwd_ws_kernel<10,8> (+ wigner_reduce_partials)", "achieved": kernels[dom]["gbs"], "peak": peak,
                         "unit": "GB/s", "frac": kernels[dom]["frac"],
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture
                         # profiles/r01_ncu_wigner_v6_summary.txt (3 410.0 MB + 18.8 MB at 2^20 samples per launch)
                         "traffic": 3428.85e6 * (micro / 1048576.0), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes[dom] * micro},
            "pipeline_roofline": {"bytes_per_sample": 6656, "achieved_gbs": round(value / world * 6656 / 1e9, 1),
                                  "frac": round(value / world * 6656 / 1e9 / peak, 4)},
            "kernels": kernels,
            "loss": loss_value, "e2e_loss": e2e_loss,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            sample = CPU_SAMPLE
            times = cpu_reference_run(sample, 3, 1, threads)
            best = min(times)
            line["cpu_baseline"] = {"value": sample / best, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d samples (bounded sample of the workload), best of 3 after 1 warm-up, torch CPU fp32" % sample}
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
    ap.add_argument("--e2e-micro", type=int, default=1 << 20, help="micro-batch of the autograd-API end-to-end pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain launches instead of replaying a CUDA graph of the rank-local step")
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
