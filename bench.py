#!/usr/bin/env python
"""bench.py - headline benchmark of the homography-warp hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4]

A "step" is one pass of the hot path over one batch of synthetic pairs: 8-basis weights ->
corner offsets -> 8x8 DLT -> per-pixel homography flow -> bidirectional bilinear warp + validity
mask -> masked L1 -> backward to both images and to the basis weights (cfg 2, the config the
metric is quoted on; --workload cfg4 = 3x512x512 pairs from 4-pt offsets).  Output pixels of both
warp directions are the unit: value = Gpix/s over all ranks.

Multi-GPU (torchrun, one rank per GPU): the batch is sharded, every rank runs the same per-GPU
batch (weak scaling); the only exchange is the scalar-loss all-reduce (NCCL), inside the step.
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

import torch  # noqa: E402

METRIC = "Gpix/s homography DLT+flow+warp fwd/bwd at 1/2/4/8 B200; % HBM peak"
UNIT = "Gpix/s"
N_SETS = 4  # rotating input sets so that no step finds its inputs in the 126 MB L2

WORKLOADS = {
    # per-GPU batch; algorithmic bytes / output pixel / direction for train fwd+bwd = 24C+1
    "cfg2": dict(B=64, C=1, h=320, w=576, param="basis8->corner offsets->DLT", bytes_per_px=25,
                 desc="cfg2: B=64 pairs 1x320x576, 8-basis flow -> DLT -> bidirectional S1 warp + M1 mask + masked L1, fwd+bwd"),
    "cfg4": dict(B=512, C=3, h=512, w=512, param="4pt offsets->DLT", bytes_per_px=73,
                 desc="cfg4 shard: 512 pairs 3x512x512 per GPU (4096 pairs / 8), 4-pt H -> bidirectional S1 warp + M1 + L1, fwd+bwd"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=16, help="pairs per CPU-baseline / reference-arm step")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--one-op", action="store_true",
                    help="cfg2 step through the one-op ops.basis_warp_loss() (forked stream branches) instead of "
                         "basis_homography() + warp_loss(); measured equal (169.2 vs 169.0 us per step)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 - 0.15 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(workload, sample):
    """Returns (fn, pixels_per_call, description).  Executes oracle/ - allowed here only."""
    from dmhomo_b200 import synth
    from oracle import port

    wl = WORKLOADS[workload]
    C, h, w = wl["C"], wl["h"], wl["w"]
    gen = synth.generator()
    img1 = synth.noise_images(sample, C, h, w, gen).requires_grad_(True)
    img2 = synth.noise_images(sample, C, h, w, gen).requires_grad_(True)
    if workload == "cfg2":
        basis = port.gen_basis(h, w).reshape(1, 8, -1)
        wf = synth.basis_weights(sample, gen).requires_grad_(True)
        wb = synth.basis_weights(sample, gen).requires_grad_(True)

        def fn():
            for t in (img1, img2, wf, wb):
                t.grad = None
            return port.pipeline_basis(img1, img2, basis, wf, wb, variant="dlt", backward=True)["loss"].item()
    else:
        off_f = synth.corner_offsets(sample, 32.0, gen).requires_grad_(True)
        off_b = synth.corner_offsets(sample, 32.0, gen).requires_grad_(True)

        def fn():
            for t in (img1, img2, off_f, off_b):
                t.grad = None
            return port.pipeline_h4pt(img1, img2, off_f, off_b, backward=True)["loss"].item()

    return fn, 2 * sample * h * w, f"{sample} of {wl['B']} pairs of {workload} per step, fwd+bwd, torch-CPU port of the reference"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    fn, px, what = cpu_step_fn(args.workload, args.cpu_sample)
    for _ in range(max(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    val = px * args.steps / dt / 1e9
    wl = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": what},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": what},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class PairStep:
    """Static buffers + the step of one workload on one GPU."""

    def __init__(self, workload, dev, rank, two_calls=False):
        from dmhomo_b200 import ops, synth
        from dmhomo_b200.compat import hem_utils

        self.ops = ops
        self.two_calls = two_calls   # cfg2 through basis_homography() + warp_loss() instead of basis_warp_loss()
        wl = WORKLOADS[workload]
        self.wl, self.workload, self.dev = wl, workload, dev
        B, C, h, w = wl["B"], wl["C"], wl["h"], wl["w"]
        self.B, self.C, self.h, self.w = B, C, h, w
        gen = torch.Generator(device=dev).manual_seed(synth.SEED + rank)
        self.sets = []
        for _ in range(N_SETS):
            img1 = torch.rand(B, C, h, w, generator=gen, device=dev).requires_grad_(True)
            img2 = torch.rand(B, C, h, w, generator=gen, device=dev).requires_grad_(True)
            if workload == "cfg2":
                par = tuple(((torch.rand(B, 8, generator=gen, device=dev) * 2 - 1) * 4.0).requires_grad_(True)
                            for _ in range(2))
            else:
                par = ((torch.rand(2 * B, 4, 2, generator=gen, device=dev) * 2 - 1) * 32.0).requires_grad_(True)
            self.sets.append((img1, img2) + (par if isinstance(par, tuple) else (par,)))
        self.basis = hem_utils.gen_basis(h, w).to(dev) if workload == "cfg2" else None
        self.src2 = synth.corner_points(2 * B, h, w, dev)
        self.pixels = 2 * B * h * w
        self.loss_vec = torch.zeros(2, device=dev, dtype=torch.float64)
        self.ev = None  # (begin, end) events around the dominant kernel, set per timed step

    def forward_backward(self, k, ev=None):
        ops = self.ops
        img1, img2, *par = self.sets[k]
        B, h, w = self.B, self.h, self.w
        if self.workload == "cfg2" and not self.two_calls:
            # the whole step as one op: weights -> H || workspace zeroing, fused warp, loss finish || adjoint DLT
            ops.warp_timing_events = ev
            loss = ops.basis_warp_loss(self.basis, img1, img2, par[0], par[1])
            self.kernel_name = ops.last_warp_kernel
            ops.warp_timing_events = None
            loss.backward()
            return loss
        if self.workload == "cfg2":
            # 8 basis weights -> corner offsets -> DLT, both directions, one launch
            Hf, Hb = ops.basis_homography(self.basis, h, w, par[0], par[1])
        else:
            H = ops.dlt4(self.src2, self.src2 + par[0])
            Hf, Hb = H[:B], H[B:]
        ops.warp_timing_events = ev   # recorded tightly around the fused warp launch (no memset, no loss_finish)
        loss = ops.warp_loss([ops.WarpTerm(img2, img1, Hf), ops.WarpTerm(img1, img2, Hb)],
                             kind=ops.PARAM_HOMOGRAPHY, sampler=ops.S1, loss_form=ops.LOSS_MASKED_DIFF,
                             border_mask=True, fused=True)
        self.kernel_name = ops.last_warp_kernel
        ops.warp_timing_events = None
        loss.backward()
        return loss

    def zero_grads(self):
        for s in self.sets:
            for t in s:
                t.grad = None


def run_ours(args):
    from dmhomo_b200 import _lib, dist as ddist

    rank, local_rank, world = ddist.init("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    import torch.distributed as tdist

    # DMH_FORCE_DIST=1: exercise the multi-rank step (loss all-reduce inside the graph) on a single rank
    dist_step = world > 1
    if os.environ.get("DMH_FORCE_DIST") and world == 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        tdist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        dist_step = True

    wl = WORKLOADS[args.workload]
    st = PairStep(args.workload, dev, rank, two_calls=not args.one_op)
    K, W = args.steps, max(args.warmup, 3)
    stream = torch.cuda.Stream(dev)

    red_base = torch.tensor([0.0, float(st.B)], device=dev)
    red_scale = torch.tensor([float(st.B), 0.0], device=dev)
    red_vecs = [torch.zeros(2, device=dev) for _ in range(N_SETS)]
    comm_stream = torch.cuda.Stream(dev)
    step_done = [torch.cuda.Event() for _ in range(N_SETS)]
    comm_done = [torch.cuda.Event() for _ in range(N_SETS)]

    def reduce_loss(loss, k):
        """The path's only exchange: one all-reduce(sum) of {loss * count, count}; the global
        mean is red_vecs[k][0] / red_vecs[k][1].  One tiny kernel + one NCCL call, on a second stream behind the step's
        event so that the next step's kernels do not wait for the collective's latency (nothing downstream of the
        step consumes the reduced scalar); the compute stream waits for it before the same input set is reused, and
        the timed region ends with a full device synchronize.  Issued eagerly, never captured: NCCL collectives
        inside a CUDA graph hang on this stack (tools/dist_probe.py --graph, NCCL 2.28.9 / torch 2.11) with > 1 rank."""
        cur = torch.cuda.current_stream(dev)
        step_done[k].record(cur)
        comm_stream.wait_event(step_done[k])
        loss.record_stream(comm_stream)
        with torch.cuda.stream(comm_stream):
            torch.addcmul(red_base, red_scale, loss.detach().expand(2), out=red_vecs[k])
            tdist.all_reduce(red_vecs[k])
            comm_done[k].record(comm_stream)

    def step_eager(k, ev=None):
        st.zero_grads()
        loss = st.forward_backward(k, ev)
        if dist_step:
            reduce_loss(loss, k)
        return loss

    graphs, g_loss, g_events, launches_per_step = [], [], [], None
    with torch.cuda.stream(stream):
        # eager warm-up (also what CUDA-graph capture needs before it)
        for i in range(3):
            step_eager(i % N_SETS)
        stream.synchronize()
        n0 = _lib.launch_count()
        step_eager(0)
        stream.synchronize()
        launches_per_step = _lib.launch_count() - n0
        use_graph = not args.no_graph
        if use_graph:
            try:
                pool = None
                for k in range(N_SETS):
                    st.zero_grads()
                    g = torch.cuda.CUDAGraph()
                    evs = (torch.cuda.Event(enable_timing=True, external=True),
                           torch.cuda.Event(enable_timing=True, external=True))
                    with torch.cuda.graph(g, pool=pool, stream=stream):
                        loss = st.forward_backward(k, evs)
                    pool = g.pool()
                    graphs.append(g)
                    g_loss.append(loss)
                    g_events.append(evs)
            except Exception as e:  # capture unsupported here: fall back to eager launches (still our kernels)
                if rank == 0:
                    import traceback
                    print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
                    traceback.print_exc()
                graphs, g_loss, g_events, use_graph = [], [], [], False
                torch.cuda.synchronize()

        def run_step(i, timed_events=None):
            k = i % N_SETS
            if use_graph:
                if dist_step:
                    stream.wait_event(comm_done[k])   # the previous collective on this set's loss scalar has read it
                graphs[k].replay()
                if dist_step:
                    reduce_loss(g_loss[k], k)
                return g_loss[k]
            return step_eager(k, timed_events)

        # ---- device-resident timed region -------------------------------------------------------
        for i in range(W):
            run_step(i)
        stream.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        # barrier + synchronize immediately before the timed region: every rank starts together
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kern_ms, eager_evs = [], []
        t_wall0 = time.time()
        e0.record(stream)
        for i in range(K):
            if use_graph:
                run_step(i)
                # external events are re-recorded by every replay: read them lazily, one replay per set
            else:
                evs = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                run_step(i, evs)
                eager_evs.append(evs)
        if dist_step:
            for ev in comm_done:          # the timed region ends when the last collective has finished, too
                stream.wait_event(ev)
        e1.record(stream)
        stream.synchronize()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        t_wall1 = time.time()
        ms_total = e0.elapsed_time(e1)
        if use_graph:
            # the last replay of each set left its pair of events recorded inside the timed region
            for k in range(min(N_SETS, K)):
                try:
                    kern_ms.append(g_events[k][0].elapsed_time(g_events[k][1]))
                except Exception:
                    pass
        else:
            kern_ms = [a.elapsed_time(b) for a, b in eager_evs]
        if rank == 0:
            time.sleep(0.15)
            sampler.stop()
        clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None
        final_loss = float(run_step(0).detach())

        # ---- end-to-end: host buffers in, loss out, every step ---------------------------------------
        B, C, h, w = st.B, st.C, st.h, st.w
        host = []
        for k in range(N_SETS):
            host.append(tuple(t.detach().cpu().pin_memory() for t in st.sets[k]))
        loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in host[0])
        Ke = max(3, min(K, 20))

        # Double-buffered like a prefetching loader: the copy of step i + 1 (second stream) runs while step i computes;
        # every step still pays its own host->device copy and its own loss read-back + synchronize.  The input set a
        # copy overwrites was last read three steps earlier, and every step ends with a stream synchronize.
        copy_stream = torch.cuda.Stream(dev)
        copy_done = [torch.cuda.Event() for _ in range(N_SETS)]

        def issue_copy(i):
            k = i % N_SETS
            with torch.cuda.stream(copy_stream), torch.no_grad():
                for dst, src in zip(st.sets[k], host[k]):
                    dst.copy_(src, non_blocking=True)
                copy_done[k].record(copy_stream)

        def e2e_run(n):
            issue_copy(0)
            last = None
            for i in range(n):
                if i + 1 < n:
                    issue_copy(i + 1)
                stream.wait_event(copy_done[i % N_SETS])
                loss = run_step(i)
                loss_host.copy_(loss.detach(), non_blocking=True)
                stream.synchronize()
                last = float(loss_host)
            copy_stream.synchronize()
            return last

        e2e_run(2)
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_run(Ke)
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks: max time --------------------------------------------------------------
    times = torch.tensor([ms_total, e2e_s * 1e3, (sum(kern_ms) / len(kern_ms)) if kern_ms else 0.0], device=dev,
                         dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(times, op=tdist.ReduceOp.MAX)
    ms_total, e2e_ms, kern_avg_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        px_step_all = st.pixels * world
        value = px_step_all * K / (ms_total * 1e-3) / 1e9
        e2e_val = px_step_all * Ke / (e2e_ms * 1e-3) / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        roofline = None
        if kern_avg_ms > 0:
            alg_bytes = wl["bytes_per_px"] * st.pixels
            achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "kernel": "%s<S1,HOMOGRAPHY,FUSED,C=%d,MASKED_DIFF,dense> (both directions, one launch)" % (st.kernel_name, st.C),
                        "kernel_ms": kern_avg_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            fn, px, what = cpu_step_fn(args.workload, args.cpu_sample)
            fn()
            best = None
            t_budget = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t_budget < 10.0 and reps < 50):   # ~10 s of CPU work, at least 3 passes
                t0 = time.perf_counter()
                fn()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
                reps += 1
                if time.perf_counter() - t_budget > 30:
                    break
            cpu = {"value": px / best / 1e9, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": what + f", best of {reps} passes"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "per_gpu_batch": st.B, "global_batch": st.B * world,
                       "parallelism": f"batch-sharded x{world}", "param": wl["param"],
                       "l2": f"{N_SETS} rotating input sets ({N_SETS * 2 * st.B * st.C * st.h * st.w * 4 / 1e6:.0f} MB) > 126 MB L2",
                       "launch": "CUDA graph replay" if use_graph else "eager", "loss": final_loss,
                       "api": ("ops.basis_warp_loss" if (args.workload == "cfg2" and args.one_op) else "ops.basis_homography + ops.warp_loss" if args.workload == "cfg2" else "ops.dlt4 + ops.warp_loss")},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": Ke,
                    "pipeline": "copy of step i+1 on a second stream overlaps step i; loss read back and synchronised every step"},
            "gpu_launches": int(launches_per_step) * K, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback); "
                             "use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
