#!/usr/bin/env python
"""Benchmark of the ORBIT episodic hot path (BASELINE.json): episodes/s and query-frames/s.

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU PyTorch path (oracle port)

Workload (BASELINE.json configs[1], SURVEY.md 8d "S2"): ProtoNet + efficientnet_b0, 224x224, 5-way,
support 200 clips x 8 frames, query 80 clips x 8 frames; one STEP = one episode =
personalise(support) + predict(query) + _reset()  (2,240 frames through the extractor + the head).
Synthetic frames and a synthetic checkpoint (no network). N>1: independent episodes per rank (weak
scaling, no data-path collective); the only collective is the NCCL all-reduce of the metric counts.

Prints ONE JSON line (rank 0). `value` = device-resident inputs; `e2e` = same metric through the public
API with pinned HOST clips (H2D inside the timed region) and the logits read back to the host.
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

WORKLOAD = "S2: ProtoNet+efficientnet_b0, 224x224, 5-way, support 200 clips x 8 frames, query 80 clips x 8 frames"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='orbit_b200', choices=['orbit_b200', 'reference'])
    ap.add_argument('--gemm', type=int, default=int(os.environ.get('ORBIT_GEMM', '1')))
    ap.add_argument('--chunk', type=int, default=int(os.environ.get('ORBIT_CHUNK', '1600')))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--profile-steps', type=int, default=2)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower() == 'active':
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n):
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo', rank=rank, world_size=world,
                                device_id=torch.device(f'cuda:{local}') if torch.cuda.is_available() else None)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------------
def cpu_reference_episode_seconds(spec, sample_support, sample_query, state_dict=None, episode_index=0, threads=None):
    """Times the reference's CPU path (oracle port: plain PyTorch fp32, all host threads) on a BOUNDED sample of
    the workload -- `sample_support`/`sample_query` clips of one S2 episode -- and scales to a whole episode by
    the frame count (the head is <0.1% of the time). Returns (episode_seconds, sample_seconds, logits, info)."""
    from oracle.recogniser import OracleRecogniser
    from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    small = EpisodeSpec(spec.way, sample_support // spec.way, sample_query // spec.way, spec.clip_length, spec.frame_size)
    oracle = OracleRecogniser('efficientnet_b0', False, 'proto', spec.clip_length, 256,
                              calib_input=None if state_dict is not None else calibration_frames(spec.frame_size))
    if state_dict is not None:
        oracle.extractor.load_state_dict({k[len('feature_extractor.'):]: v.cpu() for k, v in state_dict.items()
                                          if k.startswith('feature_extractor.')}, strict=True)
    ctx, ctx_y, tgt, tgt_y = make_episode(small, index=episode_index)
    with torch.no_grad():
        oracle.extractor(ctx[0])  # warm-up (thread pool, oneDNN primitives)
    t0 = time.perf_counter()
    oracle.personalise(ctx, ctx_y)
    logits = oracle.predict(tgt)
    oracle.reset()
    dt = time.perf_counter() - t0
    frames_sample = (len(ctx) + len(tgt)) * spec.clip_length
    frames_full = spec.way * (spec.support_clips_per_class + spec.query_clips_per_class) * spec.clip_length
    info = {"cores": threads, "kind": "port",
            "sample": f"1 episode of {len(ctx)} support + {len(tgt)} query clips x {spec.clip_length} frames "
                      f"({frames_sample} frames, {dt:.1f} s) scaled x{frames_full / frames_sample:.1f} to the "
                      f"{frames_full}-frame S2 episode"}
    return dt * frames_full / frames_sample, dt, (ctx, ctx_y, tgt, tgt_y, logits), info


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port of the unmodified
    PyTorch code path; timm is not installable offline), all host threads, same metric/config."""
    from orbit_b200.synthetic import S2
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    times = []
    for i in range(args.warmup + args.steps):
        ep_s, dt, _, info = cpu_reference_episode_seconds(S2, 20, 10, None, i)
        if i >= args.warmup:
            times.append(ep_s)
    ep_s = sum(times) / len(times)
    qf = S2.way * S2.query_clips_per_class * S2.clip_length
    val = 1.0 / ep_s
    info["value"] = val
    print(json.dumps({
        "impl": "reference", "metric": "episodes_per_sec", "value": val, "unit": "episodes/s",
        "query_frames_per_sec": qf / ep_s, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ep_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "device": "cpu", "threads": info["cores"]},
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import orbit_b200
    from orbit_b200 import lib as L
    from orbit_b200.synthetic import S2, load_synthetic_checkpoint, make_episode

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device(f'cuda:{local}')
    torch.cuda.set_device(dev)
    spec = S2
    model = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', spec.clip_length, 256, False, 16)
    model._set_device(dev)
    model._send_to_device()
    model.set_test_mode(True)
    load_synthetic_checkpoint(model, spec.frame_size)
    fe = model.feature_extractor
    fe.set_option('gemm', args.gemm)
    fe.set_option('chunk_frames', args.chunk)

    # two distinct resident episodes (1.35 GB each: far larger than the 126 MB L2, so no flush is needed)
    n_res = 2
    host_eps = [make_episode(spec, index=rank * 1000 + i, pin=True) for i in range(n_res)]
    dev_eps = [(c.to(dev), cy.to(dev), t.to(dev), ty.to(dev)) for (c, cy, t, ty) in host_eps]
    qf = spec.way * spec.query_clips_per_class * spec.clip_length
    correct = torch.zeros(2, dtype=torch.int64, device=dev)   # [correct query clips, query clips]

    def step_device(i):
        c, cy, t, ty = dev_eps[i % n_res]
        model.personalise(c, cy)
        logits, am = model.predict(t, want_argmax=True)
        model._reset()
        correct[0] += (am.long() == ty).sum()
        correct[1] += ty.numel()
        return logits

    def step_host(i):
        c, cy, t, ty = host_eps[i % n_res]
        model.personalise(c, cy.to(dev, non_blocking=True))
        logits = model.predict(t)
        model._reset()
        return logits.cpu()   # the step's result is read back to the host

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier(world)
        torch.cuda.synchronize(dev)
        l0 = L.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize(dev)
        barrier(world)
        ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
        return ms, L.launches() - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    correct.zero_()
    ms, launches = timed(step_device, args.steps, max(args.warmup, 3))
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(correct)     # the run's only collective: metric counts
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel family, timed live with CUDA events on the launch stream ----------
    fe.set_option('profile', 1 if args.profile_steps else 0)
    torch.cuda.synchronize(dev)
    for i in range(args.profile_steps):
        step_device(i)
    prof = fe.profile_read()
    fe.set_option('profile', 0)
    args.profile_steps = max(1, args.profile_steps)
    total_ms = sum(p['ms'] for p in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k]['ms'])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    p = prof[dom]
    achieved = p['bytes'] / (p['ms'] * 1e-3) / 1e9 if p['ms'] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (sustained copy)" if peaks else "fallback 6650 GB/s",
                "avg_launch_us": 1e3 * p['ms'] / max(1, p['launches']),
                "algorithmic_bytes_per_launch": p['bytes'] / max(1, p['launches']),
                "achieved_tflops": p['flops'] / (p['ms'] * 1e-3) / 1e12 if p['ms'] > 0 else 0.0,
                "families": {k: {"share": v['ms'] / total_ms, "ms_per_step": v['ms'] / args.profile_steps,
                                 "GBps": (v['bytes'] / (v['ms'] * 1e-3) / 1e9) if v['ms'] > 0 else 0.0,
                                 "launches_per_step": v['launches'] // args.profile_steps}
                             for k, v in prof.items() if v['launches']}}

    # DRAM traffic of the dominant kernel family per launch, from the committed ncu capture of this same command
    # (profiles/<round>_traffic.json, written by scripts/summarise_profiles.py; ncu numbers never come from this run)
    try:
        import glob
        tf = sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_traffic.json')))[-1]
        fam = json.load(open(tf))['families'].get(dom)
        if fam:
            roofline["traffic"] = fam['dram_bytes_per_launch']
            roofline["traffic_source"] = os.path.relpath(tf, ROOT) + " (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    except Exception:
        pass

    # ---- end to end through the public API with pinned host clips -----------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 12))
        e_ms, _ = timed(step_host, e2e_steps, 2)
        ctx, _, tgt, _ = host_eps[0]
        e2e = {"value": world / (e_ms / e2e_steps / 1e3), "unit": "episodes/s",
               "h2d_bytes_per_step": int((ctx.numel() + tgt.numel()) * 4 + len(ctx) * 8),
               "d2h_bytes_per_step": int(len(tgt) * spec.way * 4), "ms_per_step": e_ms / e2e_steps,
               "query_frames_per_sec": world * qf / (e_ms / e2e_steps / 1e3)}

    # ---- CPU baseline + parity gate (rank 0, N=1 only): the oracle is the checker and the timed CPU arm -----
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ep_s, dt, (ctx, ctx_y, tgt, tgt_y, ref_logits), info = cpu_reference_episode_seconds(
            spec, 20, 10, model.state_dict(), episode_index=7)
        info["value"] = 1.0 / ep_s
        info["unit"] = "episodes/s"
        cpu = info
        model.personalise(ctx.to(dev), ctx_y.to(dev))
        logits, am = model.predict(tgt.to(dev), want_argmax=True)
        model._reset()
        parity = {"max_abs_logit_diff": float((logits.cpu() - ref_logits).abs().max()),
                  "max_abs_logit": float(ref_logits.abs().max()),
                  "argmax_equal": bool(torch.equal(am.cpu().long(), ref_logits.argmax(1))), "tolerance": 1e-3}

    if rank == 0:
        out = {
            "metric": "episodes_per_sec", "value": value, "unit": "episodes/s",
            "query_frames_per_sec": value * qf, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.gemm == 0 else ("tf32x3" if args.gemm == 1 else "tf32"), "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": 2240, "chunk_frames": args.chunk, "gemm": args.gemm,
                       "l2": "inputs 1.35 GB/episode per GPU >> 126 MB L2, two episodes alternate (no flush needed)",
                       "parallelism": f"episodes sharded over {world} GPU(s), metric all-reduce only"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
            "parity": parity,
            "frame_acc": {"correct": int(correct[0]), "total": int(correct[1])},
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
