#!/usr/bin/env python
"""bench.py -- IQ MSamples/s through the full two-channel AIS demodulation chain on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N>1) prints ONE JSON line on rank 0.
A "step" is one submit of one batch: B streams x N complex samples (BASELINE.json configs[1]: batch=1024 synthetic
CF32 streams @1536 kSPS, FM path = ModelStandard semantics, SURVEY.md 8d "Config 2"), inputs already resident in HBM.
  value     : whole-job complex samples/s (all ranks) with inputs resident in HBM, device-timed on the launch stream
  e2e       : the same metric through aisgpu_submit() with pinned HOST buffers + aisgpu_poll() (H2D and the frame
              D2H inside the timed region)
  roofline  : front-end kernel (reads every input byte): 8 B/sample x samples per launch / its CUDA-event duration,
              against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the reference's own CPU implementation (oracle/_ref fast build, shipping flags) on this box's cores
`--impl reference` times that CPU implementation alone on the same workload shape.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ais-catcher_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))  # aissynth: the seeded stimulus generator shared with the parity tests

FS = 1536000
N_CHUNK = 131072
BATCH = 1024
RESIDENT = 8
UNIQUE = 32
ALGO_BYTES_PER_SAMPLE = 8  # CF32 in; SURVEY.md 8d


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Exactly ONE line may reach stdout (the JSON result): libraries that print there (NCCL's version banner, ...) are
# diverted to stderr by pointing fd 1 at fd 2 for the lifetime of the process; emit() writes to the saved stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def make_unique_streams(n_unique, n_samples, seed0=0):
    import numpy as np
    import aissynth
    out = np.empty((n_unique, n_samples), dtype=np.complex64)
    for u in range(n_unique):
        out[u] = aissynth.random_stream(FS, n_samples, seed0 + u)[0]
    return out


def cpu_reference_rate(uniq, model, n_threads, streams_per_thread, chunks):
    """Times the reference's CPU implementation (one model instance per stream, one OS thread per core: the
    reference's own concurrency model, Device/FileRAW.cpp:205-206) on the same synthetic streams."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    fast = O.have_ref(fast=True)
    kind = "reference" if fast else "port"
    Model = (lambda **kw: O.RefModel(fast=True, **kw)) if fast else O.PortModel
    omodel = {0: O.MODEL_STANDARD, 1: O.MODEL_BASE, 2: O.MODEL_DEFAULT}[model]
    models = [[Model(model=omodel, sample_rate=FS) for _ in range(streams_per_thread)] for _ in range(n_threads)]
    nmsg = [0] * n_threads

    def work(t):
        for j, m in enumerate(models[t]):
            x = uniq[(t * streams_per_thread + j) % len(uniq)]
            nres = len(x) // N_CHUNK
            for c in range(chunks):
                cc = c % nres
                m.push(x[cc * N_CHUNK:(cc + 1) * N_CHUNK])
            nmsg[t] += m.msg_count()

    ths = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    samples = n_threads * streams_per_thread * chunks * N_CHUNK
    return samples / dt, kind, dt, sum(nmsg), samples


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    uniq = make_unique_streams(8, N_CHUNK * 2)
    T = host_threads()
    spt = 8
    # warm-up + timed steps: every step = T threads x spt streams x 1 chunk of the configs[1] workload
    rates = []
    for i in range(args.warmup + args.steps):
        r, kind, dt, nm, samples = cpu_reference_rate(uniq, args.model, T, spt, 1)
        if i >= args.warmup:
            rates.append((samples, dt))
    tot_s = sum(s for s, _ in rates)
    tot_t = sum(t for _, t in rates)
    v = tot_s / tot_t / 1e6
    line = {
        "impl": "reference", "metric": "IQ MSamples/s through full 2-ch demod chain", "value": v, "unit": "MSamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(rates)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "model": args.model, "sample_rate": FS, "chunk_samples": N_CHUNK,
                   "sample": "%d threads x %d streams x 1 chunk per step" % (T, spt)},
        "cpu_baseline": {"value": v, "unit": "MSamples/s", "cores": T, "kind": kind,
                         "sample": "%d threads x %d streams x %d samples per step, %d steps" % (T, spt, N_CHUNK, len(rates))},
        "e2e": {"value": v, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_name(args):
    m = {0: "ModelStandard (FM path)", 1: "ModelBase (FM+PLL)", 2: "ModelDefault (coherent PhaseSearchEMA)"}[args.model]
    return "batch=%d synthetic CF32 IQ streams @%d kSPS, %s, chunk %d samples/stream/launch, 1 GPU slice per rank" % (
        args.batch, FS // 1000, m, N_CHUNK)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--model", type=int, default=0, help="0 ModelStandard (FM path, configs[1]), 2 ModelDefault, 1 ModelBase")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--also-default", action="store_true", help="also time ModelDefault on the same data")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import aisgpu
    import shard

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, N, R = args.batch, N_CHUNK, RESIDENT
    t_gen = time.time()
    uniq = make_unique_streams(UNIQUE, N * R, seed0=1000 * rank)
    log("[rank %d] generated %d unique streams x %d samples in %.1fs" % (rank, UNIQUE, N * R, time.time() - t_gen))
    uniq_pin = torch.from_numpy(uniq.view(np.float32)).pin_memory()
    uniq_dev = torch.view_as_complex(uniq_pin.to(dev, non_blocking=True).view(UNIQUE, N * R, 2))
    # resident input: R chunks of [B][N]; stream b = unique[b % U] + its own AWGN realisation (so all streams differ)
    x = torch.empty((R, B, N), dtype=torch.complex64, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    for b0 in range(0, B, UNIQUE):
        nb = min(UNIQUE, B - b0)
        blk = uniq_dev[:nb].view(nb, R, N).permute(1, 0, 2)
        x[:, b0:b0 + nb, :] = blk
    noise = torch.empty((B, N), dtype=torch.complex64, device=dev)
    for r in range(R):
        torch.view_as_real(noise).normal_(0.0, 0.005, generator=g)
        x[r] += noise
    del noise
    torch.cuda.synchronize()

    eng = aisgpu.Engine(model=args.model, sample_rate=FS, n_streams=B, max_chunk=N, device=local_rank, max_frames=1 << 20)
    est = torch.cuda.ExternalStream(eng.cuda_stream(), device=dev)

    def step(i):
        eng.submit_device(x[i % R].data_ptr(), N, N)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    eng.sync()
    eng.poll()
    c0 = eng.counters()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(est)
    for i in range(args.steps):
        step(args.warmup + i)
    eng.join()
    e1.record(est)
    e1.synchronize()
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = eng.last_launches() * args.steps
    fe_times = eng.frontend_times(args.steps)
    msgs = eng.poll()
    c1 = eng.counters()
    # the only collectives of the job: MAX of the device time, SUM of a few counters (NCCL over NVLink)
    ms_max = shard.max_over_ranks(ms, device=dev)
    delta = [a - b for a, b in zip(c1, c0)]
    delta[2] = B * N * args.steps  # samples of this rank's slice (the engine counts samples per stream)
    tot = shard.gather_counts(delta, device=dev)
    total_samples = float(tot[2])
    value = total_samples / (ms_max * 1e-3) / 1e6
    msgs_per_s = float(tot[1]) / (ms_max * 1e-3)

    # ---- end to end: pinned host buffers -> aisgpu_submit (H2D inside) -> aisgpu_poll (frame D2H inside) ----
    host = [torch.empty((B, N, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    host[0].copy_(torch.view_as_real(x[0]).cpu())
    host[1].copy_(torch.view_as_real(x[1]).cpu())
    eng.submit_ptr(host[0].data_ptr(), N)
    eng.poll()
    barrier()
    d2h = 0
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        eng.submit_ptr(host[i & 1].data_ptr(), N)
        nfr, _ = eng.poll_count()  # sync + D2H of the frame ring + host NMEA tail; frames stay in the C structs
        d2h += 4 + 184 * nfr
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    e2e_value = world * B * N * args.e2e_steps / shard.max_over_ranks(e2e_dt, device=dev) / 1e6

    also = None
    if args.also_default and args.model != 2:
        eng2 = aisgpu.Engine(model=2, sample_rate=FS, n_streams=B, max_chunk=N, device=local_rank, max_frames=1 << 20)
        est2 = torch.cuda.ExternalStream(eng2.cuda_stream(), device=dev)
        for i in range(3):
            eng2.submit_device(x[i % R].data_ptr(), N, N)
        eng2.sync()
        eng2.poll()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks = max(4, args.steps // 2)
        a0.record(est2)
        for i in range(ks):
            eng2.submit_device(x[(3 + i) % R].data_ptr(), N, N)
        eng2.join()
        a1.record(est2)
        a1.synchronize()
        ms2 = a0.elapsed_time(a1)
        fe2 = eng2.frontend_times(ks)
        m2 = len(eng2.poll())
        also = {"workload": "same data, ModelDefault (coherent PhaseSearchEMA)", "value": B * N * ks / (ms2 * 1e-3) / 1e6,
                "unit": "MSamples/s (this rank)", "ms_per_step": ms2 / ks, "frontend_ms": sum(fe2) / len(fe2), "msgs": m2}
        eng2.close()

    # the same kernel timed alone (a sync after every submit), for the record next to the live number
    iso = []
    for i in range(6):
        eng.submit_device(x[i % R].data_ptr(), N, N)
        eng.sync()
        iso.append(eng.last_frontend_ms())
    eng.poll()
    if rank == 0:
        peak, peak_src = peaks()
        fe_ms = sum(fe_times) / max(1, len(fe_times))
        achieved = ALGO_BYTES_PER_SAMPLE * B * N / (fe_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "frontend_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        cpu = None
        if not args.no_cpu:
            T = host_threads()
            nchunks = R * 12
            rate, kind, dt, nm, samples = cpu_reference_rate(uniq, args.model, T, 2, nchunks)
            cpu = {"value": rate / 1e6, "unit": "MSamples/s", "cores": T, "kind": kind,
                   "sample": "%d threads x 2 streams x %d chunks of %d samples (%.1f s wall, %.0f core-s), %s flags" % (
                       T, nchunks, N, dt, dt * T, "-O3 -ffast-math (reference shipping)" if kind == "reference" else "strict C port")}
        line = {
            "metric": "IQ MSamples/s through full 2-ch demod chain", "value": value, "unit": "MSamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "model": args.model, "sample_rate": FS, "batch_per_gpu": B,
                       "chunk_samples": N, "resident_chunks": R, "bytes_per_step_per_gpu": B * N * 8,
                       "l2": "inputs_larger_than_L2 (1.07 GB per step vs 126 MB L2)", "parallelism": "streams sharded, %d rank(s)" % world},
            "msgs_per_s": msgs_per_s,
            "e2e": {"value": e2e_value, "unit": "MSamples/s", "h2d_bytes_per_step": B * N * 8,
                    "d2h_bytes_per_step": d2h // max(1, args.e2e_steps), "steps": args.e2e_steps},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_frontend", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "frontend_ms_per_launch": fe_ms, "frontend_share_of_step": fe_ms / (ms_max / args.steps),
                         "isolated_ms_per_launch": min(iso), "isolated_frac": ALGO_BYTES_PER_SAMPLE * B * N / (min(iso) * 1e-3) / 1e9 / peak,
                         "note": "achieved = 8 B/sample x samples per launch / live CUDA-event duration of k_frontend while the back end of the previous submit shares the GPU; isolated_* = the same launch alone",
                         "whole_chain_frac": ALGO_BYTES_PER_SAMPLE * B * N / (ms / args.steps * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu,
        }
        if also:
            line["also"] = also
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
