#!/usr/bin/env python
"""bench.py -- chunks/sec (gen + light + mesh) of the B200 chunk pipeline.

Workload (BASELINE.json configs[2], the config the metric is quoted on): ONE 128x128 region of
chunk columns ("areas") centred on the spawn area, world name "test".  A step generates the region
plus the one-area ring its border faces need (130x130 = 16 900 areas: noise -> block IDs -> trees ->
column heights) and meshes the 16 384 inner areas (face culling + vertex/index/record emission).
With N GPUs the region is SHARDED: rank r owns the rows shard_rows(128, N, r) and regenerates the
ring around its strip (strong scaling, no data-path collective: areas are pure functions of seed
and coordinates, DESIGN.md "Multi-GPU").  The same line also carries, measured in the same run:

  weak          every rank its own 128x128 region (N > 1 only; round 1's figure)
  value_cold    the strong step with a region that moves every step (no list caches, unload + load)
  config5       BASELINE.json configs[4]: 512x512 world pre-generation (generate -> save files ->
                pinned host memory), rows sharded over the ranks
  e2e           the step through the C ABI with host buffers: compact transport (save files + 4-byte
                records) as `value`; raw streams, two pipelined contexts and the C++ host mirror of
                Renderer::load_all_blocking next to it
  frame / codec / edit_storm (N = 1): the callers either side of the path

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  the CPU restatement of the reference
                                                           (oracle/, all host threads; the Rust
                                                           binary cannot be built in this image)
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REGION = 128           # areas per side of the config-3 region
C5_REGION = 512        # areas per side of the config-5 world
C5_BATCH_ROWS = 16     # config 5 is swept in batches of 16 rows x 512 areas
CPU_SAMPLE = 64        # cpu_baseline sample of the GPU arm: 64x64 areas of the same region
WORLD_NAME = "test"    # the reference's own test seed (generator.rs:157)

# algorithmic work per unit (DESIGN.md "Roofline accounting"; pinned by tests/test_op_counts.py)
F2_OPS = 58            # f64 add/mul issued per 2-D simplex octave incl. the Fbm accumulate
F3_OPS = 83            # f64 add/mul issued per 3-D simplex octave incl. the Fbm accumulate
AREA_BYTES = 32768 + 256


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_threads() -> int:
    """Every core this process may run on -- NOT OMP_NUM_THREADS (torchrun exports it as 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons during the timed region (pynvml, 4 ms period: the region lasts ~30 ms)."""

    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.004)

    def start(self):
        if self._nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_pass(seed, n_side, faithful=True, threads=0):
    """One gen+light+mesh pass of the CPU restatement over the centred n_side^2 part of the workload
    region (plus its ring); n_side = REGION is the whole config.  faithful: permutation tables
    rebuilt per area (generator.rs:51), gen parallel over areas on all threads (rayon par_iter,
    world_persistence.rs:90-93), mesh serial on one thread (renderer.rs:468-472).  generous: tables
    hoisted, mesh area-parallel too."""
    import oracle
    from voxel_game_b200.world import region_xy, spawn_region

    threads = threads or host_threads()
    ax0, ay0, _, _ = spawn_region(n_side)
    ring = region_xy(ax0 - 1, ay0 - 1, n_side + 2, n_side + 2)
    inner = region_xy(ax0, ay0, n_side, n_side)
    t0 = time.perf_counter()
    vox, mh, _, used = oracle.generate_areas(seed, ring, rebuild_tables_per_area=faithful, threads=threads)
    t1 = time.perf_counter()
    world = oracle.World(ax0 - 1, ay0 - 1, n_side + 2, n_side + 2, vox, mh)
    t2 = time.perf_counter()
    faces, recs = world.mesh_areas(inner, threads=1 if faithful else threads)
    t3 = time.perf_counter()
    return {"areas": n_side * n_side, "gen_s": t1 - t0, "mesh_s": t3 - t2,
            "total_s": (t1 - t0) + (t3 - t2), "threads": used, "faces": faces}


def cpu_pregen_pass(seed, n_side, threads=0):
    """Config 5 on the CPU: generate (faithful) + store_all_blocking (bincode + LZ4 per area on the
    rayon pool, world_persistence.rs:54-58) over a centred n_side^2 sample of the 512x512 world."""
    import oracle
    from voxel_game_b200.world import region_xy, spawn_region

    threads = threads or host_threads()
    ax0, ay0, _, _ = spawn_region(n_side)
    xy = region_xy(ax0, ay0, n_side, n_side)
    t0 = time.perf_counter()
    vox, _, _, used = oracle.generate_areas(seed, xy, rebuild_tables_per_area=True, threads=threads)
    t1 = time.perf_counter()
    total, _ = oracle.area_files_encode(vox, threads=threads)
    t2 = time.perf_counter()
    return {"areas": len(xy), "gen_s": t1 - t0, "store_s": t2 - t1, "total_s": t2 - t0, "threads": used,
            "file_bytes": int(total)}


def run_reference(args, rank, world):
    """The reference arm: the same config-3 region (all 128x128 areas + ring per step) on every host
    thread of the box.  Rank 0 alone works; the thread count is set explicitly (torchrun exports
    OMP_NUM_THREADS=1) and reported as `cores`."""
    if rank != 0:
        return 0
    import oracle

    native = oracle.use_native_build()  # -march=native on the machine that is timed
    threads = host_threads()
    seed = oracle.hash_world_name(WORLD_NAME)
    n_side = args.ref_side
    for _ in range(args.warmup):
        cpu_pass(seed, n_side, threads=threads)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_pass(seed, n_side, threads=threads)
    dt = time.perf_counter() - t0
    value = n_side * n_side * args.steps / dt
    gen = cpu_pass(seed, n_side, faithful=False, threads=threads)
    c5 = cpu_pregen_pass(seed, 64, threads=threads)
    whole = n_side == REGION
    sample = (f"{'the whole' if whole else f'the centred {n_side}x{n_side} of the'} {REGION}x{REGION} region per step (+ ring), "
              f"gen parallel on {last['threads']} threads with per-area table rebuild ({last['gen_s']:.2f} s), "
              f"mesh on 1 thread as the reference ({last['mesh_s']:.2f} s)")
    line = {
        "impl": "reference", "metric": "chunks/sec (gen+light+mesh)", "value": value, "unit": "chunks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"c3_{REGION}x{REGION}_gen_light_mesh", "world_name": WORLD_NAME,
                   "seed": seed, "sample": sample, "areas_per_step": n_side * n_side,
                   "oracle_build": "-O3 -march=native -ffp-contract=off" if native else "-O3 -march=x86-64-v2 -ffp-contract=off (no gcc on this box)",
                   "note": "C restatement of the reference's CPU path (oracle/); the Rust/rayon binary cannot be built in this image (no cargo)"},
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": last["threads"], "kind": "port",
                         "sample": sample,
                         "generous_value": gen["areas"] / gen["total_s"],
                         "generous_note": "tables hoisted per seed, mesh area-parallel on all threads"},
        "config5": {"value": c5["areas"] / c5["total_s"], "unit": "chunks/s", "cores": c5["threads"],
                    "sample": f"64x64 areas of the {C5_REGION}x{C5_REGION} world: generate {c5['gen_s']:.2f} s + "
                              f"store_all_blocking (bincode + LZ4 on all threads) {c5['store_s']:.2f} s",
                    "file_bytes_per_area": c5["file_bytes"] / c5["areas"]},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-side", type=int, default=REGION,
                    help="reference arm: side of the centred part of the region per step (default: all of it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip frame / codec / edit storm / config 5 / e2e variants")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    import voxel_game_b200 as vg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:  # keep each rank (and the pinned host buffers it first-touches) next to its GPU;
        try:       # not at N = 1, where the CPU baseline must see every host core
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created; stdout must carry
        # exactly one JSON line, so it goes to stderr for the duration of the set-up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    seed = vg.hash_world_name(WORLD_NAME)
    ctx = vg.Context(local_rank, seed=seed)
    # a stream of our own: torch's default stream is the NULL stream, which vx_set_stream reads as
    # "use the context's own stream" -- the events below must sit on the stream the kernels run on
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    plans = {}

    def plan(shift_x):
        """Rows of the region (moved by shift_x areas) per rank.  One rank: all of them.  Several: strips of
        equal COST, not equal height -- a cave-zone column costs ~30x an ordinary one, and the caves are
        not spread evenly: vx_cave_columns (a cost hint from the seed alone) weighs the rows, every rank
        computes the same cut (voxel_game_b200/world.py balanced_rows)."""
        if shift_x not in plans:
            ax0, ay0, nx, ny = vg.spawn_region(REGION)
            if world == 1:
                plans[shift_x] = [(0, ny)]
            else:
                cc, cv = ctx.cave_columns(vg.region_xy(ax0 + shift_x - 1, ay0 - 1, nx + 2, ny + 2))
                plans[shift_x] = vg.balanced_rows(*vg.row_costs(cc, cv, nx, ny), world)
        return plans[shift_x]

    def strip(shift_x=0, whole=False, plan_of=None):
        """(generate list, mesh list) of this rank: its rows of the region + the ring around them."""
        ax0, ay0, nx, ny = vg.spawn_region(REGION)
        ax0 += shift_x
        lo, hi = (0, ny) if whole else plan(shift_x if plan_of is None else plan_of)[rank]
        return (vg.region_xy(ax0 - 1, ay0 + lo - 1, nx + 2, hi - lo + 2), vg.region_xy(ax0, ay0 + lo, nx, hi - lo))

    def timed(fn, steps, warmup, sample_clocks=False, finish=None):
        """W untimed + K timed calls of fn(i) between barriers and device syncs; CUDA events on the
        stream the kernels run on; -> (ms for the K steps, max over ranks; clocks).  `finish` runs after
        the K calls, inside the timed region (deferred copies of the last step)."""
        for i in range(warmup):
            fn(i)
        if finish:
            finish()
        ctx.timings()  # clears the accumulators: kernel times cover the timed steps only
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        if finish:
            finish()
        e1.record(stream)
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), clocks

    # ---- device-resident throughput (`value`): the sharded region, same lists every step
    gen_xy, mesh_xy = strip()
    n_gen, n_mesh = len(gen_xy), len(mesh_xy)
    last = {}

    def step(_i):
        ctx.generate(gen_xy, read=False)  # enqueued; vx_mesh is ordered behind it and waits for the totals
        last["info"] = ctx.mesh(mesh_xy)

    ctx.set_timing(True)
    ms, clocks = timed(step, args.steps, args.warmup, sample_clocks=True)
    info = last["info"]
    value = REGION * REGION * args.steps / (ms * 1e-3)
    t = ctx.timings()  # kernel times (CUDA events around every launch) accumulated over the timed steps
    kern = {k: t[k] / args.steps for k in ("gen_columns_ms", "gen_caves_ms", "gen_finish_ms", "mesh_count_ms", "mesh_emit_ms")}
    launches = t["mesh_launches"] + t["gen_launches"]
    work = {k: t[k] for k in ("simplex2_evals", "simplex3_evals", "cave_voxels", "cave_columns", "cave_evals")}
    faces_total = int(sum_over_ranks(info["n_faces"]))
    recs_total = int(sum_over_ranks(info["n_recs"]))
    gen_total = int(sum_over_ranks(n_gen))

    # ---- the same step with cold lists (`value_cold`): the region moves by 130 areas every step and
    # the previous one is unloaded, so no call ever sees a list or a resident set it has seen before
    ctx.set_timing(False)
    cold_steps = args.steps
    # (the strips keep the main region's row cut, so every step has the same shape: this measures the
    # step without caches, not a re-planning per step)
    cold_lists = [strip(shift_x=130 * (i + 1), plan_of=0) for i in range(cold_steps + 3)]
    prev = [gen_xy]

    def cold_step(i):
        g, m = cold_lists[i]
        ctx.unload(prev[0])
        ctx.generate(g, read=False)
        ctx.mesh(m)
        prev[0] = g

    cold_ms, _ = timed(cold_step, cold_steps, 3)
    ctx.unload(prev[0])
    value_cold = REGION * REGION * cold_steps / (cold_ms * 1e-3)

    # ---- weak scaling (N > 1): every rank its own whole region, as in round 1
    weak = None
    if world > 1:
        wg, wm = strip(shift_x=rank * REGION, whole=True)

        def weak_step(_i):
            ctx.generate(wg, read=False)
            ctx.mesh(wm)

        w_ms, _ = timed(weak_step, args.steps, 3)
        weak = {"value": len(wm) * world * args.steps / (w_ms * 1e-3), "unit": "chunks/s", "scaling": "weak",
                "ms_per_step": w_ms / args.steps, "areas_meshed_per_gpu": len(wm),
                "note": "every rank generates and meshes its own 128x128 region (round 1's bench line)"}
        ctx.unload(wg)
    ctx.generate(gen_xy, read=False)
    info = ctx.mesh(mesh_xy)  # the resident mesh the sections below read

    extras = not args.no_extras
    # ---- per-frame visibility over the resident mesh (K7; not part of `value`): a camera on the
    # spawn surface with a render distance that covers the whole region
    frame = None
    if world == 1 and extras:
        view = vg.make_view((8.0, 8.0, 58.0), (9.0, 8.5, 59.0), render_size=REGION // 2)
        ctx.set_timing(True)
        for _ in range(3):
            vis = ctx.visible(view, read=False)
        ctx.timings()  # clears the accumulators
        f_ms = 0.0
        for _ in range(10):
            vis = ctx.visible(view, read=False)
            f_ms += ctx.timings()["visible_ms"]
        f_ms /= 10
        _, _, flags = ctx.visible(view)
        rf = np.empty(n_mesh + 1, np.uint32)
        ctx.mesh_read(info, out=(rf, None, None, None, None))
        recs_in_view = int((np.diff(rf.astype(np.int64)) * flags).sum())
        frame_bytes = recs_in_view * 16 + vis["n_visible"] * 4
        frame = {"ms": f_ms, "visible_areas": vis["n_areas"], "records_tested": recs_in_view,
                 "visible_voxels": vis["n_visible"], "visible_faces": vis["n_faces"],
                 "records_per_s": recs_in_view / (f_ms * 1e-3), "algorithmic_bytes": frame_bytes}

    # ---- L-ext (north star only; the reference has no voxel light): the max(neighbour - 1) relaxation
    # over the 32x32 areas around the spawn, levels 0..15 per voxel
    light = None
    if world == 1 and extras:
        lx0, ly0, _, _ = vg.spawn_region(32)
        lxy = vg.region_xy(lx0, ly0, 32, 32)
        ctx.set_timing(True)
        ctx.light(lxy)
        ctx.timings()
        levels, l_iters = ctx.light(lxy)
        l_ms = ctx.timings()["light_ms"]
        light = {"areas": len(lxy), "ms": l_ms, "launches": l_iters + 2, "chunks_per_s": len(lxy) / (l_ms * 1e-3),
                 "lit_voxels": int((levels > 0).sum()),
                 "algorithmic_bytes": len(lxy) * (32768 * 2 + 256 + (l_iters) * (3 * 32768 + 4 * 2048)),
                 "note": "extension, no reference parity (oracle/light_ext.c is the spec): init + one relaxation launch "
                         "per round until no area changes; each launch relaxes every area to its own fixed point in "
                         "shared memory with a one-voxel halo"}

    # ---- save files (K8 / K9; not part of `value`): kernel times of the encoder and the decoder
    codec = None
    if world == 1 and extras:
        ctx.set_timing(True)
        total = ctx.encode_areas(gen_xy, read=False)  # first call: sizes the buffers (and loads the kernels)
        ctx.timings()
        blob, offs = ctx.encode_areas(gen_xy)
        enc_ms = ctx.timings()["codec_ms"]
        with vg.Context(local_rank, seed=seed) as other:
            other.set_timing(True)
            other.decode_areas(gen_xy, blob, offs, read_heights=False)
            other.timings()  # clears the accumulators: the first call also loaded the kernels
            other.decode_areas(gen_xy, blob, offs, read_heights=False)
            dec_ms = other.timings()["codec_ms"]
        codec = {"encode_ms": enc_ms, "decode_ms": dec_ms, "areas": n_gen, "file_bytes_total": int(total),
                 "file_bytes_per_area": total / n_gen, "ratio_vs_raw": n_gen * 32768 / total}
        del blob

    # ---- config 5 (BASELINE.json configs[4]): 512x512 world pre-generation, rows sharded over the
    # ranks, swept in batches of 16 rows: generate -> compress on the device -> files and heights in
    # pinned host memory (the bytes store_all_blocking writes) -> unload.  Two contexts on two host
    # threads take alternate batches, so the copies of one overlap the kernels of the other.
    config5 = None
    if extras:
        cx0, cy0, cnx, cny = vg.spawn_region(C5_REGION)
        lo, hi = vg.shard_rows(cny, world, rank)
        batches = [vg.region_xy(cx0, cy0 + r0, cnx, min(C5_BATCH_ROWS, hi - r0)) for r0 in range(lo, hi, C5_BATCH_ROWS)]
        max_areas = max(len(b) for b in batches)
        workers = [ctx, vg.Context(local_rank, seed=seed)]
        pins = [{"files": vg.PinnedArray((max_areas * 2600 + 4096,), np.uint8),
                 "off": vg.PinnedArray((max_areas + 1,), np.uint64),
                 "mh": vg.PinnedArray((max_areas, 256), np.uint8)} for _ in range(2 * len(workers))]
        c5_bytes = [0, 0]

        def c5_worker(k, count_bytes):
            c, sets = workers[k], pins[2 * k: 2 * k + 2]
            c.set_deferred_reads(True)  # the files of a batch cross to the host while the next one is generated
            for j, b in enumerate(batches[k::2]):
                p = sets[j & 1]
                c.generate(b, max_height_out=p["mh"].array[: len(b)], read=False)
                total = c.encode_areas(b, read=False)
                c.wait_reads()  # the other set is complete (this is where a writer thread would take it)
                c.encode_read(len(b), total, out=p["files"].array, offsets_out=p["off"].array[: len(b) + 1])
                c.unload(b)
                if count_bytes:
                    c5_bytes[k] += int(total)
            c.wait_reads()
            c.set_deferred_reads(False)

        def c5_sweep(count_bytes=False):
            th = [threading.Thread(target=c5_worker, args=(k, count_bytes)) for k in range(2)]
            for x in th:
                x.start()
            for x in th:
                x.join()

        workers[1].set_timing(False)
        ctx.unload(gen_xy)
        c5_sweep()  # warm-up (sizes pools and codec buffers)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c5_sweep(count_bytes=True)
        torch.cuda.synchronize()
        c5_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        c5_files = sum_over_ranks(float(sum(c5_bytes)))
        c5_d2h = c5_files + (8 + 256) * C5_REGION * C5_REGION
        config5 = {"value": C5_REGION * C5_REGION / c5_s, "unit": "chunks/s", "ms": c5_s * 1e3,
                   "areas": C5_REGION * C5_REGION, "areas_per_gpu": (hi - lo) * cnx, "batch_areas": max_areas,
                   "d2h_bytes": int(c5_d2h), "file_bytes_per_area": c5_files / (C5_REGION * C5_REGION),
                   "timing": "host wall clock around the whole sweep (device idle before and after), max over ranks",
                   "note": "vx_generate + vx_encode_areas + vx_encode_read (deferred: the files cross while the next batch is "
                           "generated) into pinned host memory + vx_unload per batch; "
                           "two contexts on two host threads per GPU take alternate batches; parity: "
                           "tests/test_gpu_parity.py::test_config5_512x512_pregeneration_sample"}
        workers[1].close()
        for p in pins:
            for a in p.values():
                a.free()
        ctx.generate(gen_xy, read=False)
        info = ctx.mesh(mesh_xy)

    # ---- edit storm (BASELINE.json configs[3]; not part of `value`): 100 000 place/break events over a
    # loaded 64x64 region in batches of 1 000 -- vx_apply_edits (World::set + column heights), then
    # vx_update_locations (Renderer::update_location: <= 7 voxel meshes per edit) read back to the host
    edit_storm = None
    if world == 1 and extras:
        with vg.Context(local_rank, seed=seed) as ec:
            n_side = 64
            ex0, ey0, _, _ = vg.spawn_region(n_side)
            exy = vg.region_xy(ex0, ey0, n_side, n_side)
            ec.set_mesh_retention(True)
            ec.generate(vg.region_xy(ex0 - 1, ey0 - 1, n_side + 2, n_side + 2), read=False)
            ec.mesh(exy)
            vox0, _ = ec.read_areas(exy)
            rng = np.random.default_rng(0x5EEDED17)
            n_edits, batch = 100_000, 1000
            gx = rng.integers(0, n_side * 16, n_edits)
            gy = rng.integers(0, n_side * 16, n_edits)
            gz = rng.integers(1, 127, n_edits)
            area = (gx // 16) + n_side * (gy // 16)  # region_xy order: index = ix + nx * iy
            cur = vox0[area, (gx % 16) + 16 * (gy % 16) + 256 * gz]
            palette = np.array([1, 6, 8, 9, 11, 26], np.uint32)  # Cobblestone Brick Boards Stone Lamp Glass
            edits = np.zeros(n_edits, vg.EDIT_DTYPE)
            edits["gx"] = ex0 * 16 + gx
            edits["gy"] = ey0 * 16 + gy
            edits["gz"] = gz
            edits["voxel"] = np.where(cur != 0, 0, palette[rng.integers(0, 6, n_edits)])
            locs = np.stack([edits["gx"], edits["gy"], edits["gz"]], axis=1).astype(np.uint32)

            def storm(lo_, hi_, incremental):
                recs_n = faces_n = dirty_n = 0
                for b0 in range(lo_, hi_, batch):
                    if incremental:
                        ec.apply_edits(edits[b0:b0 + batch], want_dirty=False)
                        di, _r, _v, _i = ec.update_locations(locs[b0:b0 + batch])
                        recs_n += di["n_recs"]
                        faces_n += di["n_faces"]
                    else:
                        d = ec.apply_edits(edits[b0:b0 + batch])
                        faces_n += ec.mesh(d)["n_faces"]
                        dirty_n += len(d)
                ec.sync()
                return recs_n, faces_n, dirty_n

            # the water simulator's tick on the device: every lake surface voxel of the region
            # whose sides are loaded, as one ordered check list -- a sequential chain by construction
            a_i, l_i = np.nonzero(vox0 == 14)  # WaterSource
            wx = ex0 * 16 + (a_i % n_side) * 16 + (l_i & 15)
            wy = ey0 * 16 + (a_i // n_side) * 16 + ((l_i >> 4) & 15)
            wz = l_i >> 8
            inside = (wx > ex0 * 16) & (wx < (ex0 + n_side) * 16 - 1) & (wy > ey0 * 16) & (wy < (ey0 + n_side) * 16 - 1)
            wcheck = np.stack([wx, wy, wz], axis=1)[inside][:4000].astype(np.uint32)
            water = None
            if len(wcheck):
                ec.water_tick(wcheck[:64])
                t0 = time.perf_counter()
                wchanged, wnext = ec.water_tick(wcheck)
                wwall = time.perf_counter() - t0
                water = {"locations": int(len(wcheck)), "ms": wwall * 1e3, "us_per_location": wwall * 1e6 / len(wcheck),
                         "changed": int(len(wchanged)), "queued_for_next_tick": int(len(wnext)),
                         "note": "vx_water_tick over the lake-surface sources of the region in array order: one warp "
                                 "applies the reference's loop body location by location (the order is an input), its "
                                 "lanes fetching the seven voxels side by side; latency-bound by construction"}

            # the per-frame pass over the RESIDENT meshes of a render zone (mesh retention: nothing is re-meshed
            # when the zone slides or an edit lands): render distance 31 around the spawn = 63 x 63 areas
            zone = vg.render_zone(vg.world.SPAWN_AREA, 31)
            zview = vg.make_view((8.0, 8.0, 58.0), (9.0, 8.5, 59.0), render_size=31)
            ec.set_timing(True)
            for _ in range(3):
                zinfo = ec.visible_zone(zone, zview, read=False)
            ec.timings()
            z_ms = 0.0
            for _ in range(10):
                zinfo = ec.visible_zone(zone, zview, read=False)
                z_ms += ec.timings()["visible_ms"]
            ec.set_timing(False)
            zone_frame = {"ms": z_ms / 10, "zone_areas": int(len(zone)), "visible_areas": zinfo["n_areas"],
                          "visible_voxels": zinfo["n_visible"], "visible_faces": zinfo["n_faces"],
                          "note": "vx_visible_zone over the retained face-mask volumes of the zone (32 KiB per visible area)"}

            half = n_edits // 2
            storm(0, 3 * batch, True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r_n, f_n, _ = storm(3 * batch, half, True)
            inc_wall = time.perf_counter() - t0
            inc_edits = half - 3 * batch
            storm(half, half + 3 * batch, False)
            t0 = time.perf_counter()
            _, f2_n, d_n = storm(half + 3 * batch, n_edits, False)
            re_wall = time.perf_counter() - t0
            re_edits = n_edits - half - 3 * batch
            edit_storm = {
                "edits": n_edits, "batch": batch, "region": "64x64 areas",
                "incremental": {"ms_per_batch": inc_wall * 1e3 / (inc_edits // batch), "edits_per_s": inc_edits / inc_wall,
                                "voxel_meshes_per_edit": r_n / inc_edits, "faces_per_edit": f_n / inc_edits,
                                "d2h_bytes_per_edit": (r_n * 16 + f_n * 172) / inc_edits,
                                "note": "end to end, wall clock: vx_apply_edits + vx_update_locations + vx_update_read "
                                        "(records, vertices, indices of the <= 7 voxels per edit in host memory)"},
                "water_tick": water,
                "zone_frame": zone_frame,
                "remesh_dirty_areas": {"ms_per_batch": re_wall * 1e3 / (re_edits // batch), "edits_per_s": re_edits / re_wall,
                                       "areas_per_batch": d_n / (re_edits // batch), "faces_per_edit": f2_n / re_edits,
                                       "note": "round 1's form: vx_apply_edits + vx_mesh of the dirty areas, meshes stay on the device"},
            }

    # ---- end to end through the C ABI with host buffers (`e2e`)
    # The pinned buffers are placed when they are allocated: do that from a core next to the GPU
    # (at N = 1 the process is otherwise free to roam, the CPU baseline needs every core), or the
    # copies cross the socket link and the number depends on where the scheduler had put us.
    saved_affinity = os.sched_getaffinity(0)
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    ctx.set_timing(False)
    files_total = ctx.encode_areas(gen_xy, read=False)

    def compact_buffers():
        p = {"files": vg.PinnedArray((int(files_total * 1.1) + 4096,), np.uint8),
             "mh": vg.PinnedArray((n_gen, 256), np.uint8),
             "rf": vg.PinnedArray((n_mesh + 1,), np.uint32),
             "packed": vg.PinnedArray((int(info["n_recs"] * 1.05) + 1024,), np.uint32)}
        for a in p.values():
            a.array.view(np.uint8).reshape(-1)[::4096] = 0  # touch every page while still next to the GPU
        return p

    def compact_step(c, p):
        """Everything the host needs of a step, in its compact form: the areas as the reference's own
        save files (LZ4 of the block IDs, what load_blocking reads), the column heights, and the mesh
        as 4-byte records (the 16-byte record, vertices and indices are a pure function of them)."""
        c.generate(gen_xy, max_height_out=p["mh"].array, read=False)
        total = c.encode_areas(gen_xy, read=False)  # returns when the sizes are known
        inf = c.mesh(mesh_xy)                       # the vertex kernel runs ...
        c.encode_read(n_gen, total, out=p["files"].array)  # ... while the files cross to the host
        c.mesh_read_compact(inf, out=(p["rf"].array, p["packed"].array))

    pin_c = compact_buffers()
    pin_d = compact_buffers()
    for p_ in (pin_c, pin_d):
        p_["off"] = vg.PinnedArray((n_gen + 1,), np.uint64)
    e2e_steps = max(3, args.steps)

    def deferred_step(i, c=None, sets=None):
        """The same step with deferred reads (vx_set_deferred_reads): the files and the records of step i
        cross to the host while step i + 1 is generated; two sets of pinned buffers, the older one is
        complete (vx_wait_reads) before the next reads are issued."""
        c = c or ctx
        p = (sets or (pin_c, pin_d))[i & 1]
        c.generate(gen_xy, max_height_out=p["mh"].array, read=False)
        total = c.encode_areas(gen_xy, read=False)
        inf = c.mesh(mesh_xy)
        c.wait_reads()
        c.encode_read(n_gen, total, out=p["files"].array, offsets_out=p["off"].array)
        c.mesh_read_compact(inf, out=(p["rf"].array, p["packed"].array))

    ctx.set_deferred_reads(True)
    e2e_ms, _ = timed(deferred_step, e2e_steps, 3, finish=ctx.wait_reads)
    ctx.set_deferred_reads(False)
    blocking_ms, _ = timed(lambda _i: compact_step(ctx, pin_c), e2e_steps, 3)
    e2e_value = REGION * REGION * e2e_steps / (e2e_ms * 1e-3)
    h2d = sum_over_ranks((2 * n_gen + n_mesh) * 8)  # the area lists of the three calls {ax, ay}
    d2h = sum_over_ranks(files_total + 8 * (n_gen + 1) + 256 * n_gen + 4 * (n_mesh + 1) + 4 * info["n_recs"])
    e2e = {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "transport": "compact, deferred reads",
           "note": "one context, one host thread: vx_generate (heights to the host) + vx_encode_areas/_read (the areas as the "
                   "reference's save files) + vx_mesh + vx_mesh_read_compact (4-byte records) into pinned host buffers; with "
                   "vx_set_deferred_reads the copies of step i run while step i + 1 is generated (two buffer sets, vx_wait_reads "
                   "before a set is reused; the last step's copies are waited for inside the timed region)",
           "blocking": {"value": REGION * REGION * e2e_steps / (blocking_ms * 1e-3), "unit": "chunks/s",
                        "ms_per_step": blocking_ms / e2e_steps, "steps": e2e_steps,
                        "note": "the same calls, every read waited for before the next call is issued (round 2's earlier figure)"}}
    if extras:
        # (b) two contexts on two host threads, each running whole steps: copies of one overlap kernels of the other
        c2 = vg.Context(local_rank, seed=seed)
        c2.generate(gen_xy, read=False)
        c2.mesh(mesh_xy)
        c2.encode_areas(gen_xy, read=False)
        pin_c2, pin_d2 = compact_buffers(), compact_buffers()
        for p_ in (pin_c2, pin_d2):
            p_["off"] = vg.PinnedArray((n_gen + 1,), np.uint64)
        pair_steps = max(2, e2e_steps // 2)

        def pair_worker(c, sets, k):
            c.set_deferred_reads(True)
            for i in range(k):
                deferred_step(i, c, sets)
            c.wait_reads()
            c.set_deferred_reads(False)
            c.sync()

        def pair_run(k):
            th = [threading.Thread(target=pair_worker, args=(c, p, k)) for c, p in ((ctx, (pin_c, pin_d)), (c2, (pin_c2, pin_d2)))]
            for x in th:
                x.start()
            for x in th:
                x.join()

        pair_run(2)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pair_run(pair_steps)
        torch.cuda.synchronize()
        pair_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e["two_contexts"] = {"value": REGION * REGION * 2 * pair_steps / pair_s, "unit": "chunks/s",
                               "ms_per_step": pair_s * 1e3 / (2 * pair_steps), "steps": 2 * pair_steps,
                               "note": "the same deferred-read step from two host threads with one context each (wall clock)"}
        c2.close()
        for a in list(pin_c2.values()) + list(pin_d2.values()):
            a.free()

    for a in pin_d.values():
        a.free()
    if extras:
        # (c) raw streams, round 1's e2e: block IDs, heights, 16-byte records, vertices, indices
        pin = {
            "vox": vg.PinnedArray((n_gen, 32768), np.uint8), "mh": vg.PinnedArray((n_gen, 256), np.uint8),
            "rf": vg.PinnedArray((n_mesh + 1,), np.uint32), "ff": vg.PinnedArray((n_mesh + 1,), np.uint32),
            "recs": vg.PinnedArray((info["n_recs"],), vg.REC_DTYPE),
            "verts": vg.PinnedArray((info["n_faces"] * 4,), vg.VERTEX_DTYPE),
            "idx": vg.PinnedArray((info["n_faces"] * 6,), np.uint16),
        }
        mesh_out = (pin["rf"].array, pin["ff"].array, pin["recs"].array, pin["verts"].array, pin["idx"].array)
        for p_ in pin.values():
            p_.array.view(np.uint8).reshape(-1)[::4096] = 0

        def raw_step(_i):
            ctx.generate(gen_xy, voxels_out=pin["vox"].array, max_height_out=pin["mh"].array)
            inf = ctx.mesh(mesh_xy)
            ctx.mesh_read(inf, out=mesh_out)

        raw_steps = 3
        raw_ms, _ = timed(raw_step, raw_steps, 1)
        raw_d2h = sum_over_ranks(n_gen * AREA_BYTES + info["rec_bytes"] + info["vertex_bytes"] + info["index_bytes"] + 8 * (n_mesh + 1))
        e2e["raw"] = {"value": REGION * REGION * raw_steps / (raw_ms * 1e-3), "unit": "chunks/s",
                      "ms_per_step": raw_ms / raw_steps, "d2h_bytes_per_step": int(raw_d2h), "steps": raw_steps,
                      "note": "vx_generate + vx_mesh + vx_mesh_read: raw block IDs, heights, records, vertices, indices (PCIe-bound)"}
        for p_ in pin.values():
            p_.free()
    if world == 1:
        os.sched_setaffinity(0, saved_affinity)
    if extras and world == 1:
        # (d) what the game would feel: Renderer::load_all_blocking of the C++ host mirror -- device mesh,
        # transfer, and the per-voxel Mesh maps (HashMap<AreaLocation, RenderArea>) rebuilt on the host
        hr = vg.HostRenderer(ctx)
        e2e["host_renderer"] = {}
        for transport in ("compact", "raw"):
            hr.clear()
            t0 = time.perf_counter()
            hr.load_all_blocking(mesh_xy, transport=transport, threads=0)
            wall = time.perf_counter() - t0
            e2e["host_renderer"][transport] = {"value": n_mesh / wall, "unit": "chunks/s (mesh only)", "ms": wall * 1e3,
                                               "voxel_meshes": hr.voxel_count()}
        e2e["host_renderer"]["note"] = ("Renderer::load_all_blocking of host/voxel_game.hpp over the 16 384 resident areas, wall clock: "
                                        "vx_mesh + transfer + one (u8, Voxel, Mesh{Vec<Vertex>, Vec<u16>}) per visible voxel inserted into "
                                        "per-area hash maps on all host threads (renderer.rs:45-73,197-219)")
        hr.clear()
        hr.close()

    # ---- rooflines (per kernel, from the CUDA-event durations collected in the timed region)
    hbm_peak, hbm_src = measured_peaks()
    fp64_gops, _ = ctx.fp64_rate()
    s2, s3, cave_vox = work["simplex2_evals"], work["simplex3_evals"], work["cave_voxels"]
    cave3 = work["cave_evals"]  # cave octaves actually evaluated (early exit is exact)
    alt3 = s3 - cave3
    mesh_bytes = n_mesh * 9216 + info["n_faces"] * 172 + info["n_recs"] * 17
    count_bytes = n_mesh * 9216  # own S/T planes (8 KiB) + one-voxel halo of four neighbours (4 x 2048 bits)

    def fp64_roof(ops, ms_k):
        a = ops / (ms_k * 1e-3) / 1e12
        return {"bound": "fp64", "achieved": a, "peak": fp64_gops / 1e3, "unit": "TFLOP/s",
                "frac": a / (fp64_gops / 1e3), "traffic": None, "ms": ms_k,
                "peak_source": "DADD+DMUL issue rate (no FMA) measured live by vx_diag_fp64_rate"}

    def hbm_roof(nbytes, ms_k):
        a = nbytes / (ms_k * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                "traffic": None, "ms": ms_k, "peak_source": hbm_src}

    rooflines = {
        "gen_columns": fp64_roof(s2 * F2_OPS + alt3 * F3_OPS, kern["gen_columns_ms"]),
        "gen_caves": fp64_roof(cave3 * F3_OPS, kern["gen_caves_ms"]),
        "gen_finish": hbm_roof(work["cave_columns"] / 256 * (32768 + 8192 + 1024), kern["gen_finish_ms"]),
        "mesh_count": hbm_roof(count_bytes, kern["mesh_count_ms"]),
        "mesh_emit": hbm_roof(mesh_bytes, kern["mesh_emit_ms"]),
    }
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(traffic_path):  # dram bytes per launch of the whole-region step (ncu --set full)
        with open(traffic_path) as f:
            for k, v in json.load(f).items():
                if k in rooflines:
                    rooflines[k]["traffic"] = v
    dominant = max(rooflines, key=lambda k: rooflines[k]["ms"])
    roofline = dict(rooflines[dominant], kernel=dominant)

    kernel_sum = sum(kern.values())
    kernel_sum_max = max_over_ranks(kernel_sum)
    line = {
        "metric": "chunks/sec (gen+light+mesh)", "value": value, "unit": "chunks/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"c3_{REGION}x{REGION}_gen_light_mesh", "areas_meshed": REGION * REGION,
                   "areas_generated": gen_total, "areas_meshed_per_gpu": n_mesh, "areas_generated_per_gpu": n_gen,
                   "world_name": WORLD_NAME, "seed": seed, "faces": faces_total, "visible_voxels": recs_total,
                   "parallelism": f"rows of the one region sharded x{world} into strips of equal cost "
                                  "(vx_cave_columns cost hint), halo ring regenerated, no collective",
                   "rows_per_rank": [list(r) for r in plan(0)],
                   "l2": "no flush needed: a step streams ~0.55 GB of block IDs and ~1.8 GB of mesh over all ranks, "
                         "per rank > 126 MB L2 at every N <= 8"},
        "clocks": clocks,
        "value_cold": {"value": value_cold, "unit": "chunks/s", "ms_per_step": cold_ms / cold_steps, "steps": cold_steps,
                       "note": "the region moves by 130 areas every step and the previous one is unloaded: "
                               "no list or residency cache can hit (vx_unload + vx_generate + vx_mesh)"},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "rooflines": rooflines,
        "kernel_ms_per_step": kern,
        "kernel_ms_sum": kernel_sum,
        "kernel_ms_sum_max_over_ranks": kernel_sum_max,
        # the stages separately (device time of their kernels; column heights and bit
        # planes are fused into generation, so LIGHT L1 has no time of its own)
        "stage_chunks_per_s_per_gpu": {
            "gen_plus_light": n_gen / ((kern["gen_columns_ms"] + kern["gen_caves_ms"] + kern["gen_finish_ms"]) * 1e-3),
            "mesh": n_mesh / ((kern["mesh_count_ms"] + kern["mesh_emit_ms"]) * 1e-3),
        },
        "work_per_step_rank0": {"simplex2_octave_evals": s2, "simplex3_octave_evals": s3, "cave_voxels": cave_vox,
                                "cave_octave_evals": cave3, "cave_octave_evals_without_early_exit": 14 * cave_vox,
                                "f64_ops": s2 * F2_OPS + s3 * F3_OPS},
    }
    if weak is not None:
        line["weak"] = weak
    if config5 is not None:
        line["config5"] = config5
    if frame is not None:
        frame["roofline"] = hbm_roof(frame["algorithmic_bytes"], frame["ms"])
        line["frame"] = frame
    if edit_storm is not None:
        line["edit_storm"] = edit_storm
    if light is not None:
        light["roofline"] = hbm_roof(light["algorithmic_bytes"], light["ms"])
        line["light_ext"] = light
    if codec is not None:
        codec["encode_roofline"] = hbm_roof(n_gen * 32768 + codec["file_bytes_total"], codec["encode_ms"])
        codec["decode_roofline"] = hbm_roof(n_gen * (32768 + 8192 + 256) + codec["file_bytes_total"], codec["decode_ms"])
        line["codec"] = codec
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle

        native = oracle.use_native_build()
        threads = host_threads()
        if frame is not None:  # the same frame on one host thread (the reference: rayon par_iter)
            recs_h = np.empty(info["n_recs"], vg.REC_DTYPE)
            rf_h = np.empty(n_mesh + 1, np.uint32)
            ctx.mesh_read(info, out=(rf_h, None, recs_h, None, None))
            oview = oracle.make_view((8.0, 8.0, 58.0), (9.0, 8.5, 59.0), render_size=REGION // 2)
            t0 = time.perf_counter()
            oracle.filter_visible(recs_h, rf_h, mesh_xy, oview)
            frame["cpu_port_ms_1_thread"] = (time.perf_counter() - t0) * 1e3

        oseed = oracle.hash_world_name(WORLD_NAME)
        c = cpu_pass(oseed, CPU_SAMPLE, faithful=True, threads=threads)
        g = cpu_pass(oseed, CPU_SAMPLE, faithful=False, threads=threads)
        line["cpu_baseline"] = {
            "value": c["areas"] / c["total_s"], "unit": "chunks/s", "cores": c["threads"], "kind": "port",
            "sample": f"the centred {CPU_SAMPLE}x{CPU_SAMPLE} areas of the same region (+ ring): gen {c['gen_s']:.2f} s on "
                      f"{c['threads']} threads with per-area table rebuild, mesh {c['mesh_s']:.2f} s on 1 thread",
            "generous_value": g["areas"] / g["total_s"],
            "generous_note": "tables hoisted per seed, mesh area-parallel on all threads",
            "oracle_build": "-march=native" if native else "-march=x86-64-v2",
            "note": "C restatement of the reference (oracle/), not the Rust/rayon binary (no cargo in this image)",
        }
    if rank == 0:
        print(json.dumps(line), flush=True)
    for p in pin_c.values():
        p.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
