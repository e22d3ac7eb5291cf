#!/usr/bin/env python3
"""bench.py -- game-cycles/sec of the batched microRTS hot path (BASELINE.json).

Headline (N=1), BASELINE configs[1]: maps/16x16/basesWorkers16x16.xml, 65536 parallel games per GPU, RandomBiasedAI self-play,
UnitTypeTable v1 + CANCEL_BOTH, 3000-cycle cap, fully observable.  One "step" advances every game by --cycles-per-step cycles
(Game.start loop body: policy x2, issueSafe x2, cycle) in ONE launch of the step kernel; finished games restart on the device.
Before the warm-up the games are spread over the phases of a match (game g is pre-advanced by (g mod 30) * 100 cycles, untimed), so a
timed window of any length sees the stationary mix of openings, mid-games and endgames.

  value     : game-cycles/s with the state resident in HBM (inputs larger than L2: 65536 x 3.6 KB = 239 MB).
  e2e       : same metric through the public API with HOST buffers every step: H2D of the restart mask + seeds from pinned memory,
              reset_masked + step, D2H of the per-game results -- all inside the timed region, as two half-batches on two streams.
              This is the lightest contract the API has (25 bytes per game and step); the heavy one is secondary.vec.
  roofline  : frac = SURVEY 8(d) algorithmic bytes (2*(32+24*U) per game-cycle) / kernel time / measured HBM peak; beside it what
              really bounds the kernel: dram_frac (bytes the kernel moved: live counter, and the ncu capture under profiles/) and
              issue_frac (warp instructions per game-cycle from the ncu capture x live rate / issue slots).
  secondary : the other BASELINE configs in short windows of the same run, each with value / roofline / cpu_baseline:
              cfg1 (8x8 self-play), cfg3 (24x24 WorkerRush vs LightRush, A*), cfg4 (32x32 partially observable MCTS playouts from
              contact roots), cfg5 (64x64 + fused observation planes of both players every cycle; cfg5_masks adds the action masks),
              vec (the JNIGridnetVecClient flow through host buffers).  `--workload X` runs one of them alone as the main line.
  cpu_baseline / --impl reference : the CPU restatement of the Java engine (oracle/, "port": no JVM in this image) on the host cores.

Multi-GPU (torchrun): games shard by rank (weak scaling, seeds offset by rank), no data-path collective; the counters are summed by
the library's one NCCL all-reduce (mrts_batch_stats_allreduce); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAP_KEY = "16x16/basesWorkers16x16"
MAX_CYCLES = 3000
CFG3_KEYS = ["24x24/basesWorkers24x24"] + ["24x24/basesWorkers24x24" + c for c in "ABCDEFGHIJKL"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="selfplay", choices=["selfplay", "cfg1", "obs", "scripted", "rollout", "vec"],
                    help="selfplay: BASELINE configs[1] (the headline line, with the other configs as `secondary`); the others run one "
                         "secondary configuration alone: cfg1 = 8x8 self-play, scripted = cfg 3, rollout = cfg 4, obs = cfg 5, vec = RL facade")
    ap.add_argument("--games", type=int, default=65536, help="games per GPU")
    ap.add_argument("--cycles-per-step", type=int, default=100)
    ap.add_argument("--map", default=MAP_KEY)
    ap.add_argument("--utt-version", type=int, default=1, choices=[1, 2, 3])
    ap.add_argument("--with-masks", action="store_true", help="obs workload: also emit both players' bit-packed action masks every step")
    ap.add_argument("--rollouts-per-game", type=int, default=64)
    ap.add_argument("--observer", type=int, default=0, help="rollout workload: player whose partially observable view is the root (-1: fully observable)")
    ap.add_argument("--roots", default="contact", choices=["contact", "thirds"],
                    help="rollout workload: contact = every root is the first state (multiple of 50 cycles) in which the observer sees an enemy; "
                         "thirds = round-1 roots at t = 0 / 500 / 1000")
    ap.add_argument("--unit-capacity", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the headline cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-stagger", action="store_true")
    ap.add_argument("--secondary-steps", type=int, default=10)
    ap.add_argument("--prewarm-seconds", type=float, default=1.5, help="untimed steps before the warm-up, until clocks have ramped")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port on the host cores (the only part of bench.py that may execute oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_threads(work, threads, seconds, what):
    cycles = [0] * threads
    deadline = time.time() + seconds
    t0 = time.time()
    ts = [threading.Thread(target=work, args=(i, cycles, deadline)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.time() - t0
    return dict(value=sum(cycles) / dt, unit="game-cycles/s", cores=threads, kind="port",
                sample="%d game-cycles of %s on %d host threads in %.1f s; C restatement of the Java engine (oracle/), not the JVM" % (sum(cycles), what, threads, dt))


def cpu_baseline(wl, map_key, seconds, utt_version=1, threads=None, observer=0):
    from microrts_b200.maps import load_maps
    from oracle import oracle as O
    maps = load_maps()
    threads = threads or (os.cpu_count() or 1)
    utt = O.Utt(utt_version, 1)

    def work(i, cycles, deadline):
        seed = 1000003 * i
        while time.time() < deadline:
            if wl in ("selfplay", "cfg1"):
                g = O.Game(utt, maps[map_key]); g.seed(seed)
                while time.time() < deadline:
                    over, _ = g.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 250, MAX_CYCLES)
                    if over or g.time >= MAX_CYCLES:
                        break
                cycles[i] += g.time
            elif wl == "scripted":
                g = O.Game(utt, maps[CFG3_KEYS[seed % 13]])
                a0, a1 = O.ScriptedAI(O.AI_WORKER_RUSH, 0), O.ScriptedAI(O.AI_LIGHT_RUSH, 0)
                while time.time() < deadline:
                    over, _ = g.run(O.AI_WORKER_RUSH, a0, O.AI_LIGHT_RUSH, a1, 100, MAX_CYCLES)
                    if over or g.time >= MAX_CYCLES:
                        break
                cycles[i] += g.time
            elif wl in ("obs", "vec"):
                g = O.Game(utt, maps[map_key]); g.seed(seed)
                while time.time() < deadline and g.time < MAX_CYCLES and not (g.gameover and g.time > 0):
                    cycles[i] += g.run_observing(O.AI_RANDOM_BIASED, O.AI_RANDOM_BIASED, 100, MAX_CYCLES)
            else:  # rollout: roots as bench builds them (first contact at a multiple of 50 cycles), playouts from the observer's view
                g = O.Game(utt, maps[map_key]); g.seed(seed)
                while time.time() < deadline and not g.gameover and g.time < MAX_CYCLES:
                    g.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 50, MAX_CYCLES)
                    if observer < 0 or (g.po_view(observer).units()[:, 1] == 1 - observer).any():
                        break
                for k in range(16):
                    if time.time() >= deadline:
                        break
                    c = g.po_view(observer) if observer >= 0 else g.clone()
                    c.seed(seed * 64 + k)
                    t0 = c.time
                    c.simulate(t0 + 100)
                    c.evaluate(0, 0, 1)
                    cycles[i] += c.time - t0
            seed += 1

    what = {"selfplay": "RandomBiasedAI self-play on %s" % map_key, "cfg1": "RandomBiasedAI self-play on %s" % map_key,
            "scripted": "WorkerRush vs LightRush (A*) on the 13 basesWorkers24x24 variants",
            "obs": "RandomBiasedAI self-play on %s with both players' observations every cycle" % map_key,
            "vec": "RandomBiasedAI self-play on %s with both players' observations every cycle" % map_key,
            "rollout": "depth-100 playouts from contact roots on %s (playout cycles only)" % map_key}[wl]
    return _cpu_threads(work, threads, seconds, what)


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = args.workload
    key = workload_map(args)
    t_per_step = max(0.25, min(20.0, 90.0 / max(1, args.steps + args.warmup)))  # the whole arm ends within about two minutes
    for _ in range(args.warmup):
        cpu_baseline(wl, key, min(1.0, t_per_step), args.utt_version, observer=args.observer)
    vals, last = [], None
    t0 = time.time()
    for _ in range(args.steps):
        last = cpu_baseline(wl, key, t_per_step, args.utt_version, observer=args.observer)
        vals.append(last["value"])
    dt = time.time() - t0
    v = sum(vals) / len(vals)
    last["value"] = v
    out = dict(metric="game_cycles_per_sec", value=v, unit="game-cycles/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=1000.0 * dt / max(1, args.steps), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32",
               data="synthetic", impl="reference",
               config=dict(workload=workload_string(args), games_per_gpu=args.games, cycles_per_step=args.cycles_per_step,
                           max_cycles=MAX_CYCLES, note="CPU arm: each step is a bounded sample of the same workload on all host threads"),
               cpu_baseline=last, e2e=dict(value=v, unit="game-cycles/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (10 ms period), with the
    recipe's `nvidia-smi -lms` line as the fallback when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.p, self.thread, self.stop_flag = index, None, None, False
        self.sm, self.mx, self.reasons = [], None, set()

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll(self, nv, h):
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in bits:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self._visible_index()), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
        return self

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(self.sm)
            return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=self.mx, samples=len(sm), reasons=sorted(self.reasons), source="nvml")
        if not self.p:
            return dict(sm_mhz=None, sm_max_mhz=None, samples=0, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, samples=len(sm), reasons=sorted(reasons), source="nvidia-smi")


# ----------------------------------------------------------------------------------------------------------------------
class Ctx:
    """Per-process plumbing: rank / device, the library's NCCL communicator (world > 1), peaks and committed ncu figures."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        self.np, self.torch, self.dist, self.args = np, torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
        torch.cuda.set_device(self.local)
        import microrts_b200 as M
        from microrts_b200 import _ffi, sharding
        self.M, self.ffi, self.sharding = M, _ffi, sharding
        self.comm = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

            def exchange(uid):  # host plumbing: rank 0's NCCL unique id reaches the other ranks through torch.distributed
                box = [uid]
                dist.broadcast_object_list(box, src=0)
                return box[0]
            self.comm = sharding.Communicator(self.rank, self.world, self.local, exchange)
        self.peaks = None
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm_peak = float(self.peaks["hbm_gbs"]) if self.peaks and "hbm_gbs" in self.peaks else 6650.0
        self.peak_source = "measured copy bandwidth (MEASURED_PEAKS.json)" if self.peaks else "fallback (B200_PROFILING.md)"
        self.ncu = {}
        try:
            self.ncu = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        self.utt = M.UnitTypeTable(args.utt_version, 1)

    def stream(self, b):
        return self.torch.cuda.ExternalStream(self.ffi.lib().mrts_batch_stream(b._h), device=self.torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_stats(self, b):
        """Counters since the last reset, summed over every rank's batch by the library's all-reduce."""
        return self.comm.all_reduce_stats(b) if self.comm else b.stats()

    def seeds(self, n):
        return self.sharding.global_seeds(0, self.rank * n, n)  # game g of the global batch always uses seed g

    def pgs(self, key):
        return self.M.maps.standard_map(key, self.utt)

    def close(self):
        if self.comm:
            self.comm.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def stagger(ctx, b, tmp, n, period=30, chunk=100, stride=1):
    """Spread the games over the phases of a match: game g is pre-advanced by ((g // stride) mod period) * chunk cycles (untimed).
    `tmp` is a second batch of the same shape and policies with auto-reset on; each round steps a copy of the whole batch and keeps
    the result only for the games that still have to move on."""
    np = ctx.np
    phase = (np.arange(n) // stride) % period
    for k in range(1, period):
        tmp.copy_games(b)
        tmp.step(chunk, MAX_CYCLES)
        b.copy_games(tmp, mask=(phase >= k).astype(np.uint8))
    tmp.sync()
    b.sync()


def timed_window(ctx, b, step, steps, warmup, prewarm_seconds=0.0):
    """W untimed steps, then exactly K steps bracketed by barrier + synchronize; per-step CUDA events on the batch's stream."""
    torch = ctx.torch
    stream = ctx.stream(b)
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < prewarm_seconds:
        step(); b.sync()
    for _ in range(warmup):
        step()
    b.sync()
    st0, io0, l0 = b.stats(), b.io_bytes(), b.launch_count
    clocks = ClockSampler(ctx.local).start()
    ctx.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.perf_counter()
    for a, z in evs:
        a.record(stream); step(); z.record(stream)
    b.sync()
    ctx.barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    st1, io1 = b.stats(), b.io_bytes()
    kernel_ms = [a.elapsed_time(z) for a, z in evs]
    wall_max, dev_max = ctx.max_over_ranks(wall, sum(kernel_ms) / 1000.0)
    return dict(wall=wall_max, dev_s=sum(kernel_ms) / 1000.0, dev_max=dev_max, kernel_ms=kernel_ms, clocks=clk, launches=b.launch_count - l0,
                stats={k: st1[k] - st0[k] for k in st1}, totals=st1, io=(io1[0] - io0[0], io1[1] - io0[1]), kernel=b.last_kernel)


def roofline(ctx, w, algorithmic_bytes, formula, ncu_key=None, note=None):
    """SURVEY 8(d) roofline of the window's dominant kernel + what really bounds it."""
    n_launch = max(1, len(w["kernel_ms"]))
    mean_ms = sum(w["kernel_ms"]) / n_launch
    achieved = algorithmic_bytes / max(w["dev_s"], 1e-12) / 1e9
    io_total = w["io"][0] + w["io"][1]
    r = dict(bound="hbm", achieved=achieved, peak=ctx.hbm_peak, unit="GB/s", frac=achieved / ctx.hbm_peak, traffic=None,
             peak_source=ctx.peak_source, kernel=w["kernel"], bytes_formula=formula, mean_launch_ms=mean_ms,
             global_io_bytes_per_launch=io_total / n_launch,
             dram_frac=io_total / max(w["dev_s"], 1e-12) / 1e9 / ctx.hbm_peak,
             dram_frac_source="bytes of game state and outputs the kernels moved in this window (library counter mrts_batch_io_bytes) / kernel time / peak")
    cap = ctx.ncu.get(ncu_key) if ncu_key else None
    if cap:
        # figures of the committed ncu capture of this kernel on this configuration (profiles/, per round): DRAM bytes per launch and
        # warp instructions per game-cycle are properties of the code and workload; the rate they are turned into is this run's
        r["traffic"] = cap.get("dram_bytes_per_launch")
        r["traffic_source"] = cap.get("source")
        if cap.get("dram_bytes_per_launch") and cap.get("games") == ctx.args.games:
            r["dram_frac_ncu"] = cap["dram_bytes_per_launch"] / (mean_ms / 1e3) / 1e9 / ctx.hbm_peak
        wi = cap.get("warp_inst_per_game_cycle")
        sm_mhz = (w["clocks"] or {}).get("sm_mhz") or (ctx.peaks or {}).get("sm_max_mhz") or 1965.0
        cyc = max(1, w["stats"]["cycles"])
        if wi:
            r["issue_frac"] = wi * cyc / max(w["dev_s"], 1e-12) / (148 * 4 * sm_mhz * 1e6)
            r["issue_frac_source"] = "%.0f warp instructions per game-cycle (ncu capture) x this window's game-cycles/s / (148 SMs x 4 issue slots x %.0f MHz)" % (wi, sm_mhz)
        for k in ("lanes_active", "issue_active_pct"):
            if k in cap:
                r[k + "_ncu"] = cap[k]
    if note:
        r["note"] = note
    return r


def workload_map(args):
    return {"selfplay": args.map, "cfg1": "8x8/basesWorkers8x8", "scripted": CFG3_KEYS[0], "rollout": "BWDistantResources32x32",
            "obs": "GardenOfWar64x64", "vec": args.map}[args.workload]


def workload_string(args):
    if args.workload in ("selfplay", "cfg1"):
        return "maps/%s.xml x %d games/GPU, RandomBiasedAI self-play (Game.start loop), UTT v%d CANCEL_BOTH, %d-cycle cap" % (
            workload_map(args), args.games, args.utt_version, MAX_CYCLES)
    return args.workload


def state_bytes(mean_units):
    return 2.0 * (32.0 + 24.0 * mean_units)


# ----------------------------------------------------------------------------------------------------------------------
# workloads.  Each returns the fields of one JSON line (value, config, roofline, ...); rank 0 prints.
# ----------------------------------------------------------------------------------------------------------------------
def run_selfplay(ctx, key, n, C, steps, warmup, prewarm, e2e, cpu_seconds, ncu_key):
    M, np, args = ctx.M, ctx.np, ctx.args
    pgs = ctx.pgs(key)
    b = M.BatchedGameState(ctx.utt, pgs, n, device=ctx.local, unit_capacity=args.unit_capacity)
    seeds = ctx.seeds(n)
    for bb in (b,):
        bb.set_policy(0, M.POLICY_RANDOM_BIASED); bb.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.reset(seeds)
    b.set_auto_reset(True)
    if not args.no_stagger:
        tmp = M.BatchedGameState(ctx.utt, pgs, n, device=ctx.local, unit_capacity=args.unit_capacity)
        tmp.set_policy(0, M.POLICY_RANDOM_BIASED); tmp.set_policy(1, M.POLICY_RANDOM_BIASED)
        tmp.set_auto_reset(True)
        stagger(ctx, b, tmp, n)
        tmp.close()
    w = timed_window(ctx, b, lambda: b.step(C, MAX_CYCLES), steps, warmup, prewarm)
    res = b.results()
    red = ctx.sum_stats(b)
    d = w["stats"]
    cycles_all = d["cycles"] * ctx.world if ctx.world > 1 else d["cycles"]
    if ctx.world > 1:  # the window's own cycles summed over ranks: totals since reset minus what every rank had before the window
        t = ctx.torch.tensor([d["cycles"], d["unit_cycles"], d["decisions"]], dtype=ctx.torch.int64, device="cuda")
        ctx.dist.all_reduce(t)
        cycles_all, ucyc_all, dec_all = t.tolist()
    else:
        ucyc_all, dec_all = d["unit_cycles"], d["decisions"]
    mean_units = d["unit_cycles"] / max(1, d["cycles"])
    bpc = state_bytes(mean_units)
    out = dict(value=cycles_all / w["wall"], ms_per_step=1000.0 * w["wall"] / max(1, steps), clocks=w["clocks"], gpu_launches=w["launches"],
               config=dict(workload="maps/%s.xml x %d games/GPU, RandomBiasedAI self-play (Game.start loop), UTT v%d CANCEL_BOTH, %d-cycle cap" % (key, n, args.utt_version, MAX_CYCLES),
                           games_per_gpu=n, cycles_per_step=C, max_cycles=MAX_CYCLES, auto_reset="on device",
                           phases="staggered: game g pre-advanced by (g mod 30)*100 cycles before the warm-up" if not args.no_stagger else "all games start at t=0",
                           l2="state (%.0f MB/GPU) larger than L2, no flush" % (n * (80 + 7 * 4 * b.cap) / 1e6),
                           mean_live_units=ucyc_all / max(1, cycles_all), decisions_per_cycle=dec_all / max(1, cycles_all), unit_capacity=b.cap),
               roofline=roofline(ctx, w, d["cycles"] * bpc, "game-cycles x 2*(32+24*U), U = %.2f mean live units" % mean_units, ncu_key,
                                 note="frac counts the SURVEY 8(d) algorithmic bytes (state read + written once per game-cycle). The kernel keeps a game in shared "
                                      "memory for all cycles of a launch, so the bytes it really moves (dram_frac) are far fewer and frac can exceed 1: state-only "
                                      "stepping is bound by instruction issue (issue_frac), not by HBM. The HBM-bound path is secondary.cfg5"),
               stats=dict(wins_p0=red["wins_p0"], wins_p1=red["wins_p1"], draws=red["draws"], games_finished=red["games_finished"],
                          game_errors=int((res[:, 3] != 0).sum()), device_time_s=w["dev_max"], wall_time_s=w["wall"], window_game_cycles=d["cycles"],
                          step_kernel_ms=[round(x, 3) for x in w["kernel_ms"]]))
    b.close()
    if e2e:
        out["e2e"] = e2e_selfplay(ctx, pgs, n, C, steps, warmup, seeds)
    if cpu_seconds > 0 and ctx.rank == 0:
        out["cpu_baseline"] = cpu_baseline("selfplay", key, cpu_seconds, args.utt_version)
    return out


def e2e_selfplay(ctx, pgs, n, C, steps, warmup, seeds):
    """The batch driven as two halves, each a BatchedGameState with its own stream: while the device steps one half, the host reads
    the other half's results, picks the games to restart and sends their mask and seeds.  Every byte crosses PCIe inside the timed
    region and every step of a half waits for its results; only the host's work and the copies of one half overlap the other's kernel."""
    M, np, torch = ctx.M, ctx.np, ctx.torch
    halves = []
    cuts = [0, n // 2, n] if n >= 2 else [0, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        hb = M.BatchedGameState(ctx.utt, pgs, hi - lo, device=ctx.local)
        hb.set_policy(0, M.POLICY_RANDOM_BIASED); hb.set_policy(1, M.POLICY_RANDOM_BIASED)
        hb.set_auto_reset(False)
        hb.reset(seeds[lo:hi])
        h = dict(b=hb, n=hi - lo, mask=torch.zeros(hi - lo, dtype=torch.uint8).pin_memory(), seeds=torch.from_numpy(seeds[lo:hi].copy()).pin_memory(),
                 res=torch.zeros((hi - lo, 4), dtype=torch.int32).pin_memory(), episode=0, before=0)
        h["mask_np"], h["seeds_np"], h["res_np"] = h["mask"].numpy(), h["seeds"].numpy(), h["res"].numpy()
        halves.append(h)

    def submit(h):
        h["before"] = int(np.where(h["mask_np"] != 0, 0, h["res_np"][:, 0]).sum())
        h["b"].reset_masked(h["mask_np"], h["seeds_np"])   # H2D: restart mask (n bytes) + seeds (8n bytes)
        h["b"].step(C, MAX_CYCLES)                         # asynchronous on the half's stream

    def collect(h):
        h["b"].results(h["res_np"])                         # D2H: per-game {time, winner, gameover, errors} (16n bytes); waits
        res_np = h["res_np"]
        done = (res_np[:, 2] != 0) | (res_np[:, 0] >= MAX_CYCLES)
        h["mask_np"][:] = done
        if done.any():
            h["episode"] += 1
            h["seeds_np"][done] += ctx.world * n * h["episode"]
        return int(res_np[:, 0].sum()) - h["before"]

    for h in halves:
        submit(h)
    for _ in range(warmup):
        for h in halves:
            collect(h); submit(h)
    for h in halves:
        collect(h)
    ctx.barrier()
    t0 = time.perf_counter()
    adv = 0
    for h in halves:
        submit(h)
    for k in range(steps):
        for h in halves:
            adv += collect(h)
            if k + 1 < steps:
                submit(h)
    ctx.barrier()
    e_wall = time.perf_counter() - t0
    for h in halves:
        h["b"].close()
    (e_wall,) = ctx.max_over_ranks(e_wall)
    ea = torch.tensor([adv], dtype=torch.int64, device="cuda")
    if ctx.world > 1:
        ctx.dist.all_reduce(ea)
    return dict(value=ea.item() / e_wall, unit="game-cycles/s", h2d_bytes_per_step=9 * n, d2h_bytes_per_step=16 * n, ms_per_step=1000.0 * e_wall / steps,
                how="two half-batches on two streams: the host work and copies of one overlap the kernel of the other",
                note="lightest contract of the API: per game and step 9 bytes in (restart mask + seed) and 16 bytes out (time, winner, gameover, "
                     "errors); states, observations and actions stay on the device. The RL-style contract with observations, masks and actions "
                     "crossing PCIe every cycle is secondary.vec")


def run_scripted(ctx, n, C, steps, warmup, prewarm, cpu_seconds):
    M, args = ctx.M, ctx.args
    pgs = [ctx.pgs(k) for k in CFG3_KEYS]

    def make():
        bb = M.BatchedGameState(ctx.utt, pgs, n, device=ctx.local, scripted_ai=True)
        bb.set_policy(0, M.POLICY_WORKER_RUSH, M.PF_ASTAR); bb.set_policy(1, M.POLICY_LIGHT_RUSH, M.PF_ASTAR)
        bb.set_auto_reset(True)
        return bb
    b = make()
    b.reset(ctx.seeds(n))
    b.set_auto_reset(True)
    if not args.no_stagger:
        tmp = make()
        stagger(ctx, b, tmp, n, period=12, chunk=100, stride=13)  # the 13 variants of one phase sit next to each other
        tmp.close()
    w = timed_window(ctx, b, lambda: b.step(C, MAX_CYCLES), steps, warmup, prewarm)
    red = ctx.sum_stats(b)
    d = w["stats"]
    cycles_all = sum_over_ranks(ctx, d["cycles"])
    mean_units = d["unit_cycles"] / max(1, d["cycles"])
    out = dict(value=cycles_all / w["wall"], ms_per_step=1000.0 * w["wall"] / max(1, steps), clocks=w["clocks"], gpu_launches=w["launches"],
               config=dict(workload="maps/24x24/basesWorkers24x24{,A..L}.xml (13 variants round-robin) x %d games/GPU, WorkerRush vs LightRush with A*, %d-cycle cap "
                                    "(deterministic: 13 distinct games, phases staggered by 100 cycles)" % (n, MAX_CYCLES),
                           games_per_gpu=n, cycles_per_step=C, max_cycles=MAX_CYCLES, mean_live_units=mean_units, unit_capacity=b.cap),
               roofline=roofline(ctx, w, d["cycles"] * (state_bytes(mean_units) + 16.0 * mean_units), "game-cycles x (2*(32+24*U) + 2*8*U)", "cfg3",
                                 note="a sequential decision chain per game (A* queries inside): bound by instruction fetch/issue, nowhere near HBM"),
               stats=dict(games_finished=red["games_finished"], wins_p0=red["wins_p0"], wins_p1=red["wins_p1"], draws=red["draws"], game_errors=red["errors"],
                          device_time_s=w["dev_max"], wall_time_s=w["wall"], window_game_cycles=d["cycles"]))
    b.close()
    if cpu_seconds > 0 and ctx.rank == 0:
        out["cpu_baseline"] = cpu_baseline("scripted", CFG3_KEYS[0], cpu_seconds)
    return out


def sum_over_ranks(ctx, v):
    if ctx.world == 1:
        return v
    t = ctx.torch.tensor([int(v)], dtype=ctx.torch.int64, device="cuda")
    ctx.dist.all_reduce(t)
    return t.item()


def run_obs(ctx, n, C, steps, warmup, prewarm, with_masks, cpu_seconds):
    M, torch, args = ctx.M, ctx.torch, ctx.args
    key = "GardenOfWar64x64"
    pgs = ctx.pgs(key)

    def make():
        bb = M.BatchedGameState(ctx.utt, pgs, n, device=ctx.local, unit_capacity=args.unit_capacity)
        bb.set_policy(0, M.POLICY_RANDOM_BIASED); bb.set_policy(1, M.POLICY_RANDOM_BIASED)
        bb.set_auto_reset(True)
        return bb
    b = make()
    b.reset(ctx.seeds(n))
    b.set_auto_reset(True)
    if not args.no_stagger:
        tmp = make()
        stagger(ctx, b, tmp, n)
        tmp.close()
    obs = [torch.empty((n, 6, b.height, b.width), dtype=torch.uint8, device="cuda") for _ in range(2)]
    b.set_observation_outputs(obs[0], obs[1])
    mbytes = (b.mask_width + 7) // 8
    msk = None
    if with_masks:
        msk = [torch.empty((n, b.height, b.width, mbytes), dtype=torch.uint8, device="cuda") for _ in range(2)]
        if hasattr(b, "set_mask_outputs"):
            b.set_mask_outputs(msk[0], msk[1])

    def step():
        b.step(C, MAX_CYCLES)
        if with_masks and not hasattr(b, "set_mask_outputs"):
            b.masks(0, "bits", out=msk[0]); b.masks(1, "bits", out=msk[1])
    w = timed_window(ctx, b, step, steps, warmup, prewarm)
    red = ctx.sum_stats(b)
    d = w["stats"]
    cycles_all = sum_over_ranks(ctx, d["cycles"])
    mean_units = d["unit_cycles"] / max(1, d["cycles"])
    bytes_total = d["cycles"] * state_bytes(mean_units) + steps * n * 2 * 6 * b.height * b.width
    form = "game-cycles x 2*(32+24*U) + steps x games x 2 x 6*H*W (uint8 planes of both players)"
    name = "maps/%s.xml x %d games/GPU, RandomBiasedAI self-play, %d cycle(s) per step, 6-plane uint8 observations of BOTH players written every step (fused)" % (key, n, C)
    if with_masks:
        bytes_total += steps * n * 2 * b.height * b.width * mbytes
        form += " + steps x games x 2 x H*W*%d (bit-packed masks)" % mbytes
        name += " + both players' bit-packed action masks (%d bits = %d bytes per cell)" % (b.mask_width, mbytes)
    out = dict(value=cycles_all / w["wall"], ms_per_step=1000.0 * w["wall"] / max(1, steps), clocks=w["clocks"], gpu_launches=w["launches"],
               config=dict(workload=name, games_per_gpu=n, cycles_per_step=C, max_cycles=MAX_CYCLES, mean_live_units=mean_units, unit_capacity=b.cap,
                           l2="state + outputs (%.1f GB/GPU) larger than L2, no flush" % ((n * 2 * 6 * b.height * b.width * (1 + (mbytes / 6.0 if with_masks else 0))) / 1e9)),
               roofline=roofline(ctx, w, bytes_total, form, "cfg5_masks" if with_masks else "cfg5",
                                 note="the HBM-bound path: write-only traffic; the write-only ceiling of this part is below the copy peak (profiles/r2*_write_ceiling.json)"),
               stats=dict(games_finished=red["games_finished"], game_errors=red["errors"], device_time_s=w["dev_max"], wall_time_s=w["wall"], window_game_cycles=d["cycles"]))
    ceil = ctx.ncu.get("write_ceiling_gbs")
    if ceil:
        out["roofline"]["write_ceiling_gbs"] = ceil
        out["roofline"]["frac_of_write_ceiling"] = out["roofline"]["achieved"] / ceil
    b.close()
    if cpu_seconds > 0 and ctx.rank == 0:
        out["cpu_baseline"] = cpu_baseline("obs", key, cpu_seconds)
    return out


def build_contact_roots(ctx, roots, n, observer):
    """Every root becomes the first state (at a multiple of 50 cycles of RandomBiasedAI self-play) in which `observer` sees an enemy
    unit: found on the device through the evaluation of the observer's view (no visible enemy <=> SimpleSqrtEvaluationFunction3 == 1).
    Games that never make contact get a copy of another game's contact root.  Returns the fraction that made contact itself."""
    M, np = ctx.M, ctx.np
    play = M.BatchedGameState(ctx.utt, roots.maps[0], n, device=ctx.local)
    play.set_policy(0, M.POLICY_RANDOM_BIASED); play.set_policy(1, M.POLICY_RANDOM_BIASED)
    play.reset(ctx.seeds(n))
    frozen = np.zeros(n, dtype=bool)
    for _t in range(50, MAX_CYCLES + 1, 50):
        play.step(50, MAX_CYCLES)
        ev = play.evaluate(0, max(observer, 0), observer)
        res = play.results()
        contact = (ev != 1.0) & ~frozen & (res[:, 2] == 0)
        if contact.any():
            roots.copy_games(play, mask=contact.astype(np.uint8))
            frozen |= contact
        if frozen.all():
            break
    play.close()
    frac = float(frozen.mean())
    if not frozen.all() and frozen.any():
        src = np.nonzero(frozen)[0]
        idx = np.arange(n, dtype=np.int64)
        idx[~frozen] = src[np.arange(int((~frozen).sum())) % len(src)]
        roots.copy_games(roots, src_index=idx, mask=(~frozen).astype(np.uint8))
    roots.sync()
    return frac


def run_rollout(ctx, n, steps, warmup, prewarm, cpu_seconds):
    M, np, torch, args = ctx.M, ctx.np, ctx.torch, ctx.args
    key = "BWDistantResources32x32"
    R, observer = args.rollouts_per_game, args.observer
    b = M.BatchedGameState(ctx.utt, ctx.pgs(key), n, device=ctx.local)
    b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
    b.reset(ctx.seeds(n))
    contact_frac = None
    if args.roots == "contact" and observer >= 0:
        contact_frac = build_contact_roots(ctx, b, n, observer)
        roots_desc = "contact roots: the first state (multiple of 50 cycles of RandomBiasedAI self-play) in which player %d sees an enemy unit" % observer
    else:
        third = n // 3
        for t in range(0, 1000, 100):
            if t == 0:
                snap0 = b.export(0, third)
            if t == 500:
                snap500 = b.export(third, third)
            b.step(100, MAX_CYCLES)
        b.import_(snap0, 0); b.import_(snap500, third)
        roots_desc = "roots advanced to t = 0 / 500 / 1000 by thirds"
    nr = n * R
    ev = torch.empty(nr, dtype=torch.float32, device="cuda"); tm = torch.empty(nr, dtype=torch.int32, device="cuda")
    L = ctx.ffi.lib()

    def step():
        rc = L.mrts_batch_rollout(b._h, R, 100, 0, max(observer, 0), observer, None, ev.data_ptr(), tm.data_ptr(), 1)
        assert rc == 0, L.mrts_last_error()
    w = timed_window(ctx, b, step, steps, warmup, prewarm)
    d = w["stats"]
    cycles_all = sum_over_ranks(ctx, d["cycles"])
    rollouts = steps * nr
    root_units = float(np.mean(b.export()["header"][:, 3])) if n <= 32768 else float(np.mean(b.export(0, 32768)["header"][:, 3]))
    mean_len = d["cycles"] / max(1, rollouts)
    out = dict(value=cycles_all / w["wall"], ms_per_step=1000.0 * w["wall"] / max(1, steps), clocks=w["clocks"], gpu_launches=w["launches"],
               config=dict(workload="maps/%s.xml: %d roots/GPU x %d NaiveMCTS playouts (RandomBiasedAI both sides, depth 100) from %s, SimpleSqrtEvaluationFunction3; %s"
                                    % (key, n, R, "fully observable roots" if observer < 0 else "player %d's partially observable view" % observer, roots_desc),
                           roots_per_gpu=n, rollouts_per_step=nr, mean_rollout_cycles=mean_len, rollouts_per_s=rollouts * ctx.world / w["wall"],
                           contact_fraction=contact_frac, root_units=root_units,
                           note="game-cycles/s counts cycles actually simulated (rollouts end when the observer's view holds one side only), not rollouts x 100"),
               roofline=roofline(ctx, w, rollouts * (32.0 + 24.0 * root_units + 8.0), "rollouts x (32+24*U0+8): root read once, {eval, cycles} written once", "cfg4",
                                 note="roots are read from L2/HBM once per rollout and played in shared memory: bound by instruction issue"),
               stats=dict(device_time_s=w["dev_max"], wall_time_s=w["wall"], window_game_cycles=d["cycles"]))
    b.close()
    if cpu_seconds > 0 and ctx.rank == 0:
        out["cpu_baseline"] = cpu_baseline("rollout", key, cpu_seconds, observer=observer)
    return out


def run_vec(ctx, n_envs, steps, warmup, prewarm, cpu_seconds):
    from microrts_b200 import vec_bench
    return vec_bench.run(ctx, n_envs, steps, warmup, prewarm, cpu_seconds, cpu_baseline, timed_window, roofline)


# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    ctx = Ctx(args)
    wl = args.workload
    cpu_s = 0.0 if args.no_cpu_baseline else args.cpu_seconds
    sec_cpu = 0.0 if args.no_cpu_baseline else min(args.cpu_seconds, 4.0)
    n, C = args.games, args.cycles_per_step
    if wl == "selfplay":
        main = run_selfplay(ctx, args.map, n, C, args.steps, args.warmup, args.prewarm_seconds, not args.no_e2e, cpu_s, "cfg2" if args.map == MAP_KEY else None)
    elif wl == "cfg1":
        main = run_selfplay(ctx, "8x8/basesWorkers8x8", n, C, args.steps, args.warmup, args.prewarm_seconds, not args.no_e2e, cpu_s, "cfg1")
    elif wl == "scripted":
        main = run_scripted(ctx, n, C, args.steps, args.warmup, args.prewarm_seconds, cpu_s)
    elif wl == "obs":
        main = run_obs(ctx, n, C if C != 100 else 1, args.steps, args.warmup, args.prewarm_seconds, args.with_masks, cpu_s)
    elif wl == "rollout":
        main = run_rollout(ctx, n if n != 65536 else 16384, args.steps, args.warmup, args.prewarm_seconds, cpu_s)
    else:
        main = run_vec(ctx, n if n != 65536 else 16384, args.steps, args.warmup, args.prewarm_seconds, cpu_s)
    secondary = None
    if wl == "selfplay" and not args.no_secondary:
        K, W = max(3, min(args.secondary_steps, args.steps)), 3
        secondary = {}
        for name, fn in (("cfg1", lambda: run_selfplay(ctx, "8x8/basesWorkers8x8", n, C, K, W, 0.3, False, sec_cpu, "cfg1")),
                         ("cfg3", lambda: run_scripted(ctx, n, C, K, W, 0.3, sec_cpu)),
                         ("cfg4", lambda: run_rollout(ctx, 16384, K, W, 0.3, sec_cpu)),
                         ("cfg5", lambda: run_obs(ctx, n, 1, 2 * K, W, 0.3, False, sec_cpu)),
                         ("cfg5_masks", lambda: run_obs(ctx, n, 1, 2 * K, W, 0.3, True, 0.0)),
                         ("vec", lambda: run_vec(ctx, 16384, 2 * K, W, 0.3, 0.0))):
            try:
                r = fn()
                r.setdefault("unit", "game-cycles/s")
                r["steps"], r["warmup"] = (2 * K if name in ("cfg5", "cfg5_masks", "vec") else K), W
                secondary[name] = r
            except Exception as e:  # a secondary window must never take the headline line down with it
                secondary[name] = dict(error="%s: %s" % (type(e).__name__, e))
    if ctx.rank == 0:
        out = dict(metric="game_cycles_per_sec", value=main.pop("value"), unit="game-cycles/s", n_gpus=ctx.world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=main.pop("ms_per_step"), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32", data="synthetic",
                   config=main.pop("config"), clocks=main.pop("clocks"), e2e=main.pop("e2e", None), gpu_launches=main.pop("gpu_launches"),
                   roofline=main.pop("roofline"), cpu_baseline=main.pop("cpu_baseline", None))
        out.update(main)
        if secondary is not None:
            out["secondary"] = secondary
        print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
