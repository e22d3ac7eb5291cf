#!/usr/bin/env python3
"""bench.py -- game-cycles/sec of the batched microRTS hot path (BASELINE.json configs[1]).

Workload (N=1): maps/16x16/basesWorkers16x16.xml, 65536 parallel games per GPU, RandomBiasedAI self-play,
UnitTypeTable v1 + CANCEL_BOTH, 3000-cycle cap, fully observable.  One "step" advances every game by
--cycles-per-step cycles (Game.start loop body: policy x2, issueSafe x2, cycle) in ONE launch of the step kernel;
finished games restart on the device (auto-reset), so every timed window is a stationary mix of game phases.
With the defaults (150 steps x 100 cycles) the timed window covers five full 3000-cycle games per slot.

  value     : game-cycles/s with the state resident in HBM (inputs larger than L2: 65536 x 3.6 KB = 239 MB).
  e2e       : same metric through the public API with HOST buffers every step: H2D of the restart mask + seeds from
              pinned memory, reset_masked + step, D2H of the per-game results -- all inside the timed region.  The batch
              is driven as two half-batches on two streams so that the host's part of one overlaps the other's kernel.
  roofline  : algorithmic bytes = 2*(32+24*U) per game-cycle (SURVEY 8d, U = mean live units measured in the run)
              x game-cycles per launch / mean launch time (CUDA events on the batch's stream) vs measured HBM peak.
  cpu_baseline / --impl reference : the CPU restatement of the Java engine (oracle/, "port": no JVM in this image)
              on the host cores, bounded sample.

Multi-GPU (torchrun): games shard by rank (weak scaling, 65536 per GPU, seeds offset by rank), no data-path
collective; one NCCL all-reduce of the win/draw counters at the end; time = max over ranks.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

MAP_KEY = "16x16/basesWorkers16x16"
MAX_CYCLES = 3000


def workload_string(args):
    return "maps/%s.xml x %d games/GPU, RandomBiasedAI self-play (Game.start loop), UTT v%d CANCEL_BOTH, %d-cycle cap" % (
        args.map, args.games, args.utt_version, MAX_CYCLES)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--with-masks", action="store_true",
                    help="obs workload: also emit both players' bit-packed action masks every step (the masks variant of SURVEY 8d cfg 5)")
    ap.add_argument("--utt-version", type=int, default=1, choices=[1, 2, 3],
                    help="selfplay workload: UnitTypeTable version (1 = VERSION_ORIGINAL, the headline; 2 = VERSION_ORIGINAL_FINETUNED, "
                         "the secondary variant of SURVEY 8d; 3 = VERSION_NON_DETERMINISTIC)")
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=65536, help="games per GPU")
    ap.add_argument("--cycles-per-step", type=int, default=100)
    ap.add_argument("--map", default=MAP_KEY)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--prewarm-seconds", type=float, default=1.5, help="untimed steps before the warm-up, until clocks have ramped")
    ap.add_argument("--workload", default="selfplay", choices=["selfplay", "obs", "scripted", "rollout"],
                    help="selfplay: BASELINE configs[1] (the default, the headline line); the others are the secondary "
                         "configurations of SURVEY 8(d): obs = cfg 5 (64x64 + fused observations every cycle), scripted = cfg 3 "
                         "(24x24 WorkerRush vs LightRush, A*), rollout = cfg 4 (32x32 partially observable MCTS playouts)")
    ap.add_argument("--rollouts-per-game", type=int, default=64)
    ap.add_argument("--observer", type=int, default=0, help="rollout workload: player whose partially observable view is the root (-1: fully observable roots)")
    ap.add_argument("--unit-capacity", type=int, default=0, help="unit slots per game (0 = automatic bound; a game that needs more sets its error flag)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores (bench.py's cpu_baseline leg may execute oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_baseline(map_key, seconds, threads=None, utt_version=1):
    import golden_io
    from oracle import oracle as O
    maps = golden_io.load_maps()
    threads = threads or (os.cpu_count() or 1)
    utt = O.Utt(utt_version, 1)
    cycles = [0] * threads
    games_done = [0] * threads
    deadline = time.time() + seconds

    def work(i):
        seed = 1000003 * i
        while time.time() < deadline:
            g = O.Game(utt, maps[map_key])
            g.seed(seed)
            seed += 1
            while time.time() < deadline:
                over, _ = g.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, 250, MAX_CYCLES)
                if over or g.time >= MAX_CYCLES:
                    break
            cycles[i] += g.time
            games_done[i] += 1

    t0 = time.time()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.time() - t0
    return dict(value=sum(cycles) / dt, unit="game-cycles/s", cores=threads, kind="port",
                sample="%d RandomBiasedAI self-play games (%d game-cycles) of %s on %d host threads in %.1f s; C restatement of the Java "
                       "engine (oracle/), not the JVM" % (sum(games_done), sum(cycles), map_key, threads, dt))


def cpu_baseline_secondary(wl, seconds, n_rollouts_per_root=8, threads=None):
    """The oracle port on the host cores for the secondary workloads (bounded sample, same definition of a game-cycle)."""
    import golden_io
    from oracle import oracle as O
    maps = golden_io.load_maps()
    threads = threads or (os.cpu_count() or 1)
    utt = O.Utt(1, 1)
    cycles = [0] * threads
    deadline = time.time() + seconds

    def work(i):
        seed = 1000003 * i
        while time.time() < deadline:
            if wl == "scripted":
                keys = ["24x24/basesWorkers24x24"] + ["24x24/basesWorkers24x24" + c for c in "ABCDEFGHIJKL"]
                g = O.Game(utt, maps[keys[seed % 13]])
                a0, a1 = O.ScriptedAI(O.AI_WORKER_RUSH, 0), O.ScriptedAI(O.AI_LIGHT_RUSH, 0)
                while time.time() < deadline:
                    over, _ = g.run(O.AI_WORKER_RUSH, a0, O.AI_LIGHT_RUSH, a1, 100, MAX_CYCLES)
                    if over or g.time >= MAX_CYCLES:
                        break
                cycles[i] += g.time
            elif wl == "obs":
                g = O.Game(utt, maps["GardenOfWar64x64"]); g.seed(seed)
                while time.time() < deadline and g.time < MAX_CYCLES and not (g.gameover and g.time > 0):
                    cycles[i] += g.run_observing(O.AI_RANDOM_BIASED, O.AI_RANDOM_BIASED, 100, MAX_CYCLES)
            else:
                g = O.Game(utt, maps["BWDistantResources32x32"]); g.seed(seed)
                g.run(O.AI_RANDOM_BIASED, None, O.AI_RANDOM_BIASED, None, (seed % 3) * 500, MAX_CYCLES)
                for k in range(n_rollouts_per_root):
                    if time.time() >= deadline:
                        break
                    c = g.po_view(0) if wl == "rollout" else g.clone()
                    c.seed(seed * 64 + k)
                    t0 = c.time
                    c.simulate(t0 + 100)
                    c.evaluate(0, 0, 1)
                    cycles[i] += c.time - t0
            seed += 1

    t0 = time.time()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.time() - t0
    return dict(value=sum(cycles) / dt, unit="game-cycles/s", cores=threads, kind="port",
                sample="%d game-cycles of the same workload on %d host threads in %.1f s; C restatement of the Java engine (oracle/), not the JVM" % (sum(cycles), threads, dt))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_per_step = max(0.25, min(20.0, 90.0 / max(1, args.steps + args.warmup)))  # the whole arm ends within about two minutes
    for _ in range(args.warmup):
        cpu_baseline(args.map, min(1.0, t_per_step), utt_version=args.utt_version)
    vals = []
    t0 = time.time()
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(args.map, t_per_step, utt_version=args.utt_version)
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
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self._visible_index()), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

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


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import golden_io
    import microrts_b200 as M
    import parity as P
    maps = golden_io.load_maps()
    utt = M.UnitTypeTable(args.utt_version, 1)
    pgs = M.PhysicalGameState.fromXML(P.map_to_xml(maps[args.map]), utt)
    n, C = args.games, args.cycles_per_step
    b = M.BatchedGameState(utt, pgs, n, device=local)
    from microrts_b200 import _ffi
    stream = torch.cuda.ExternalStream(_ffi.lib().mrts_batch_stream(b._h), device=torch.device("cuda", local))
    from microrts_b200 import sharding
    seeds = sharding.global_seeds(0, rank * n, n)  # game g of the global batch always uses seed g
    b.set_policy(0, M.POLICY_RANDOM_BIASED)
    b.set_policy(1, M.POLICY_RANDOM_BIASED)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- clock ramp: untimed steps of the same kernel until the GPU has been busy for a while -------------------------
    b.reset(seeds)
    b.set_auto_reset(True)
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < args.prewarm_seconds:
        b.step(C, MAX_CYCLES)
        b.sync()

    # ---- device-resident run ("value") -------------------------------------------------------------------------------
    b.reset(seeds)
    b.set_auto_reset(True)
    for _ in range(args.warmup):
        b.step(C, MAX_CYCLES)
    b.sync()
    st0 = b.stats()
    l0 = b.launch_count
    clocks = ClockSampler(local)
    clocks.start()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.perf_counter()
    for a, z in evs:
        a.record(stream)
        b.step(C, MAX_CYCLES)
        z.record(stream)
    b.sync()
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    launches = b.launch_count - l0
    st1 = b.stats()
    kernel_ms = [a.elapsed_time(z) for a, z in evs]
    dev_s = sum(kernel_ms) / 1000.0
    cycles = st1["cycles"] - st0["cycles"]
    ucyc = st1["unit_cycles"] - st0["unit_cycles"]
    decisions = st1["decisions"] - st0["decisions"]
    res = b.results()
    errors = int((res[:, 3] != 0).sum())

    tmax = torch.tensor([wall, dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    wall_max, dev_max = tmax.tolist()
    # the single NCCL reduce of win/score statistics (counters of the timed window; wins/draws since the reset)
    red = sharding.reduce_stats(dict(st1, cycles=cycles, unit_cycles=ucyc, decisions=decisions, errors=errors), device="cuda")
    cycles_all, ucyc_all, dec_all = red["cycles"], red["unit_cycles"], red["decisions"]
    w0, w1, dr, fin, err_all = red["wins_p0"], red["wins_p1"], red["draws"], red["games_finished"], red["errors"]
    value = cycles_all / wall_max
    mean_units = ucyc_all / max(1, cycles_all)

    # ---- end-to-end through the public API with host buffers ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # The batch is driven as two halves, each a BatchedGameState with its own stream: while the device steps one half,
        # the host reads the other half's results, picks the games to restart and sends their mask and seeds.  Every
        # byte still crosses PCIe inside the timed region and every step of a half waits for its results; only the
        # host's work and the copies of one half overlap the kernel of the other.
        halves = []
        cuts = [0, n // 2, n] if n >= 2 else [0, n]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            hb = M.BatchedGameState(utt, pgs, hi - lo, device=local)
            hb.set_policy(0, M.POLICY_RANDOM_BIASED)
            hb.set_policy(1, M.POLICY_RANDOM_BIASED)
            hb.set_auto_reset(False)
            hb.reset(seeds[lo:hi])
            h = dict(b=hb, n=hi - lo,
                     mask=torch.zeros(hi - lo, dtype=torch.uint8).pin_memory(),
                     seeds=torch.from_numpy(seeds[lo:hi].copy()).pin_memory(),
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
                h["seeds_np"][done] += world * n * h["episode"]
            return int(res_np[:, 0].sum()) - h["before"]

        for h in halves:
            submit(h)
        for _ in range(args.warmup):
            for h in halves:
                collect(h)
                submit(h)
        for h in halves:
            collect(h)
        barrier()
        t0 = time.perf_counter()
        adv = 0
        for h in halves:
            submit(h)
        for k in range(args.steps):
            for h in halves:
                adv += collect(h)
                if k + 1 < args.steps:
                    submit(h)
        barrier()
        e_wall = time.perf_counter() - t0
        for h in halves:
            h["b"].close()
        et = torch.tensor([e_wall], dtype=torch.float64, device="cuda")
        ea = torch.tensor([adv], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            dist.all_reduce(ea, op=dist.ReduceOp.SUM)
        e2e = dict(value=ea.item() / et.item(), unit="game-cycles/s", h2d_bytes_per_step=9 * n, d2h_bytes_per_step=16 * n,
                   ms_per_step=1000.0 * et.item() / args.steps,
                   how="two half-batches on two streams: the host work and copies of one overlap the kernel of the other")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_step) ---------------------------------------------------------------------
    peaks, peak_src = None, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks["hbm_gbs"]) if peaks and "hbm_gbs" in peaks else 6650.0
    if peaks:
        peak_src = "measured (MEASURED_PEAKS.json)"
    bytes_per_cycle = 2.0 * (32.0 + 24.0 * (ucyc / max(1, cycles)))
    achieved = (cycles * bytes_per_cycle / max(1, len(kernel_ms))) / (dev_s / max(1, len(kernel_ms))) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_step_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    peak_source=peak_src, kernel="k_step_fast", bytes_per_game_cycle=bytes_per_cycle,
                    mean_launch_ms=sum(kernel_ms) / max(1, len(kernel_ms)),
                    note="achieved = SURVEY 8(d) algorithmic bytes (state read + written once per game-cycle) / kernel time. The kernel keeps a game in shared memory for all cycles of a launch, so the bytes it really moves (`traffic`, ncu) are ~70x fewer and frac can exceed 1: state-only stepping is instruction-issue bound, not HBM-bound (profiles/r1v_*: issue slots 78 % busy). The HBM-bound path is --workload obs")

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(args.map, args.cpu_seconds, utt_version=args.utt_version)

    out = dict(metric="game_cycles_per_sec", value=value, unit="game-cycles/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=1000.0 * wall_max / max(1, args.steps), higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="int32", data="synthetic",
               config=dict(workload=workload_string(args),
                           games_per_gpu=n, cycles_per_step=C, max_cycles=MAX_CYCLES, auto_reset="on device",
                           l2="state (%.0f MB/GPU) larger than L2, no flush" % (n * (64 + 7 * 4 * b.cap) / 1e6),
                           mean_live_units=mean_units, decisions_per_cycle=dec_all / max(1, cycles_all), unit_capacity=b.cap),
               clocks=clk, e2e=e2e, gpu_launches=launches, roofline=roofline, cpu_baseline=cpu,
               stats=dict(wins_p0=w0, wins_p1=w1, draws=dr, games_finished=fin, game_errors=err_all,
                          device_time_s=dev_max, wall_time_s=wall_max, step_kernel_ms=[round(x, 3) for x in kernel_ms]))
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# secondary workloads (SURVEY 8d cfg 3/4/5): same timing rules, one JSON line each, no CPU / e2e legs
# ----------------------------------------------------------------------------------------------------------------------
def run_secondary(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import golden_io
    import microrts_b200 as M
    import parity as P
    from microrts_b200 import _ffi, sharding
    maps = golden_io.load_maps()
    utt = M.UnitTypeTable(1, 1)
    n = args.games
    seeds = sharding.global_seeds(0, rank * n, n)
    wl = args.workload
    obs = None
    if wl == "obs":
        key = "GardenOfWar64x64"
        b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps[key]), utt), n, device=local, unit_capacity=args.unit_capacity)
        b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
        obs = [torch.empty((n, 6, b.height, b.width), dtype=torch.uint8, device="cuda") for _ in range(2)]
        b.set_observation_outputs(obs[0], obs[1])
        C = args.cycles_per_step if args.cycles_per_step != 100 else 1
        name = "maps/%s.xml x %d games/GPU, RandomBiasedAI self-play, %d cycle(s) per step, 6-plane uint8 observations of BOTH players written every step (fused)" % (key, n, C)
        if args.with_masks:
            mbytes = (b.mask_width + 7) // 8
            msk = [torch.empty((n, b.height, b.width, mbytes), dtype=torch.uint8, device="cuda") for _ in range(2)]
            name += " + both players' bit-packed action masks (%d bits = %d bytes per cell)" % (b.mask_width, mbytes)
    elif wl == "scripted":
        keys = ["24x24/basesWorkers24x24"] + ["24x24/basesWorkers24x24" + c for c in "ABCDEFGHIJKL"]
        pgs = [M.PhysicalGameState.fromXML(P.map_to_xml(maps[k]), utt) for k in keys]
        b = M.BatchedGameState(utt, pgs, n, device=local, scripted_ai=True)
        b.set_policy(0, M.POLICY_WORKER_RUSH, M.PF_ASTAR); b.set_policy(1, M.POLICY_LIGHT_RUSH, M.PF_ASTAR)
        C = args.cycles_per_step
        name = "maps/24x24/basesWorkers24x24{,A..L}.xml (13 variants round-robin) x %d games/GPU, WorkerRush vs LightRush with A*, %d-cycle cap (deterministic: 13 distinct games)" % (n, MAX_CYCLES)
    else:
        key = "BWDistantResources32x32"
        b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps[key]), utt), n, device=local)
        b.set_policy(0, M.POLICY_RANDOM_BIASED); b.set_policy(1, M.POLICY_RANDOM_BIASED)
        C = 100
        name = "maps/%s.xml: %d root states/GPU (RandomBiased self-play advanced to t=0/500/1000 by thirds) x %d NaiveMCTS playouts (RandomBiasedAI both sides, depth 100) from %s, SimpleSqrtEvaluationFunction3" % (key, n, args.rollouts_per_game, "fully observable roots" if args.observer < 0 else "player %d's partially observable view" % args.observer)
    stream = torch.cuda.ExternalStream(_ffi.lib().mrts_batch_stream(b._h), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    b.reset(seeds)
    rollout_out = None
    if wl == "rollout":
        # roots: thirds of the batch advanced to t = 0 / 500 / 1000
        third = n // 3
        tgt = np.zeros(n, dtype=np.int64); tgt[third:2 * third] = 500; tgt[2 * third:] = 1000
        for t in range(0, 1000, 100):
            # games whose target is reached are frozen by exporting/importing nothing: step only advances unfinished games,
            # so run the whole batch and restore the early roots afterwards
            if t == 0:
                snap0 = b.export(0, third)
            if t == 500:
                snap500 = b.export(third, third)
            b.step(100, 3000)
        b.import_(snap0, 0)
        b.import_(snap500, third)
        nr = n * args.rollouts_per_game
        ev = torch.empty(nr, dtype=torch.float32, device="cuda"); tm = torch.empty(nr, dtype=torch.int32, device="cuda")
        L = _ffi.lib()

        def step():
            rc = L.mrts_batch_rollout(b._h, args.rollouts_per_game, 100, 0, 0, args.observer, None, ev.data_ptr(), tm.data_ptr(), 1)
            assert rc == 0, L.mrts_last_error()
    else:
        b.set_auto_reset(True)

        def step():
            b.step(C, MAX_CYCLES)
            if wl == "obs" and args.with_masks:
                b.masks(0, "bits", out=msk[0]); b.masks(1, "bits", out=msk[1])

    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < args.prewarm_seconds:
        step(); b.sync()
    if wl != "rollout":
        b.reset(seeds)
        b.set_auto_reset(True)
    for _ in range(args.warmup):
        step()
    b.sync()
    st0, l0 = b.stats(), b.launch_count
    clocks = ClockSampler(local); clocks.start()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.perf_counter()
    for a, z in evs:
        a.record(stream); step(); z.record(stream)
    b.sync(); barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    st1 = b.stats()
    launches = b.launch_count - l0
    kernel_ms = [a.elapsed_time(z) for a, z in evs]
    dev_s = sum(kernel_ms) / 1000.0
    d = {k: st1[k] - st0[k] for k in st1}
    tmax = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    red = sharding.reduce_stats(dict(st1, cycles=d["cycles"], unit_cycles=d["unit_cycles"], decisions=d["decisions"], errors=d["errors"]), device="cuda")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cycles, ucyc = d["cycles"], d["unit_cycles"]
    mean_units = ucyc / max(1, cycles)
    peaks = None
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks["hbm_gbs"]) if peaks and "hbm_gbs" in peaks else 6650.0
    state_bytes = 2.0 * (32.0 + 24.0 * mean_units)
    extra = {}
    if wl == "obs":
        # B_state per game-cycle + B_obs = P*C*H*W bytes per emitted observation set (one per step)
        bytes_total = cycles * state_bytes + args.steps * n * 2 * 6 * b.height * b.width
        kern, form = "k_step_fast", "cycles*2*(32+24*U) + steps*games*2*6*H*W (uint8 planes of both players)"
        if args.with_masks:
            bytes_total += args.steps * n * 2 * b.height * b.width * ((b.mask_width + 7) // 8)
            kern, form = "k_step_fast + k_step(masks)", form + " + steps*games*2*H*W*10 (bit-packed masks)"
    elif wl == "scripted":
        bytes_total = cycles * (state_bytes + 16.0 * mean_units)
        kern, form = "k_step", "cycles*(2*(32+24*U) + 2*8*U)"
    else:
        rollouts = args.steps * n * args.rollouts_per_game
        root_units = float(np.mean(b.export()["header"][:, 3]))
        bytes_total = rollouts * (32.0 + 24.0 * root_units + 8.0)
        kern, form = "k_rollout", "rollouts*(32+24*U0+8)"
        extra = dict(rollouts_per_s=rollouts / tmax.item(), mean_rollout_cycles=cycles / max(1, rollouts), root_units=root_units)
    achieved = bytes_total / dev_s / 1e9
    obs_traffic = None  # dram bytes of one launch from the committed ncu capture, for the configuration it was taken on
    if wl == "obs" and not args.with_masks and n == 65536 and C == 1:
        try:
            obs_traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_step_fast_obs_dram_bytes_per_launch")
        except Exception:
            pass
    out = dict(metric="game_cycles_per_sec", value=red["cycles"] / tmax.item(), unit="game-cycles/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=1000.0 * tmax.item() / max(1, args.steps), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32", data="synthetic",
               config=dict(workload=name, games_per_gpu=n, cycles_per_step=C, max_cycles=MAX_CYCLES, mean_live_units=mean_units,
                           l2="state + outputs larger than L2, no flush", unit_capacity=b.cap, **extra),
               clocks=clk, e2e=None, gpu_launches=launches,
               roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=obs_traffic,
                             peak_source="measured (MEASURED_PEAKS.json)" if peaks else "fallback", kernel=kern, bytes_formula=form,
                             mean_launch_ms=sum(kernel_ms) / max(1, len(kernel_ms))),
               cpu_baseline=None if args.no_cpu_baseline else cpu_baseline_secondary("rollout_fo" if (wl == "rollout" and args.observer < 0) else wl, min(args.cpu_seconds, 6.0)),
               stats=dict(wins_p0=red["wins_p0"], wins_p1=red["wins_p1"], draws=red["draws"], games_finished=red["games_finished"],
                          game_errors=red["errors"], device_time_s=dev_s, wall_time_s=tmax.item()))
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload != "selfplay":
        run_secondary(a)
    else:
        run_ours(a)
