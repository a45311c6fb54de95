#!/usr/bin/env python
"""bench.py — Spartan prove time of one vPIN workload on N B200s (default: CNN network A, BASELINE.json configs[1]).

A "step" is what vPIN's Rust driver does per network after the R1CS is built
(vPIN_proof_generation/src/main.rs:14-46 -> proof_point_add.rs:39-98 and proof_point_mult.rs:39-98), for BOTH instances
of the network (point additions, point multiplications):
    SNARK::encode -> commit(vars_para) -> commit(vars_input) -> my_dense_mlpoly_commit(vars) -> row-wise combine -> my_lib_prove

  value : seconds per step with the instance, the public parameters (generator tables) and the three assignments already
          resident in HBM (Montgomery form); device-timed with CUDA events on the library's stream, max over ranks.
  e2e   : the same step through the host-buffer C ABI the reference's Rust driver would bind: Instance::new from host COO
          triples, SNARKGens::new (generator tables cached in the context after warm-up), encode, the three commitments
          from host assignments, my_lib_prove with host buffers, proof bytes back on the host.
  --impl reference : the CPU restatement of the reference prover (oracle/, kind "port": the Rust reference cannot be built
          here — no cargo, no crates) on the host cores, proving the SAME workload at full size every step (CNN A: both
          instances, ~24 s per step on 16 cores).
  N > 1 : `value` is ONE proof of the named network sharded over the N GPUs (strong scaling); the N-replicas arrangement
          (one network per rank, weak scaling) is reported beside it under `replicas`.

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spartan_prove_time_s"
UNIT = "s"
TRANSCRIPT_LABEL = b"snark_example"  # vPIN_proof_generation/src/proof_point_add.rs:83

# Wall-clock budget of the reference arm (seconds): the CPU port proves the full workload every step; if the projected run
# (warm-up + steps) would not fit, fewer steps are run and the line says so ("steps" is then what was actually run).
REF_BUDGET_S = float(os.environ.get("VPIN_REF_BUDGET_S", "780"))


def env_int(name, default):
    return int(os.environ.get(name, default))


def rank_world():
    return env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx = max(mx, float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------ workload
def make_workload(tag, replica=0):
    """synthetic witnesses of one named network; `replica` > 0 draws a different network of the same shape (other points,
    weights and sums — the N-GPU job proves N different networks, one per rank)"""
    from vpin_b200 import workloads as W
    m, n_add = W.SHAPES[tag]
    return dict(tag=tag, m=m, n_add=n_add, mult=W.synth_point_mult(m, seed=W.SEED + 2 * replica),
                add=W.synth_point_add(n_add, seed=W.SEED + 2 * replica + 1) if n_add else None, seeds=W.tape_seeds())


class InstanceState:
    """One R1CS instance of the workload with everything the two timed legs need, host and device side."""

    def __init__(self, ctx, kind, built, torch, device_built=False, high_priority=False):
        """built: what api.point_mult / api.point_addition return (host assignments) or, with device_built, what
        api.point_mult_device returns (assignments expanded on the device, already in Montgomery form: no large host buffer
        exists at all - only the resident leg can then be timed)"""
        from vpin_b200 import api
        self.ctx = ctx
        self.kind = kind
        self.device_built = device_built
        if device_built:
            self.dims, self.inst, d_para, d_input, d_vars, self.inputs, self.n = built
        else:
            self.dims, self.inst, vp, vi, v, self.inputs = built
        ctx.sync()
        t0 = time.time()
        self.gens = api.SNARKGens(ctx, *self.dims)  # the first one of a process derives the generators and builds their tables
        ctx.sync()
        self.gens_s = time.time() - t0
        dev = torch.device("cuda", torch.cuda.current_device())
        if device_built:
            self.d_assign = [d_para, d_input, d_vars]
        else:
            self.p_para, self.p_input, self.p_vars = self.inst.pad(vp), self.inst.pad(vi), self.inst.pad(v)
            self.n = len(self.p_vars) // 32
            self.coo = self.inst.export_coo(self.dims[1])
            # pinned host copies (e2e leg reads its inputs from pinned memory)
            self.h_coo = []
            for a in self.coo:
                t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory()
                t[: a.nbytes] = torch.from_numpy(a.view("u1").reshape(-1))
                self.h_coo.append((t, len(a)))
            self.h_assign = []
            for b in (self.p_para, self.p_input, self.p_vars):
                t = torch.empty(len(b), dtype=torch.uint8).pin_memory()
                t[:] = torch.frombuffer(bytearray(b), dtype=torch.uint8)
                self.h_assign.append(t)
            # HBM-resident Montgomery copies (value leg)
            self.d_assign = []
            for t in self.h_assign:
                d = t.to(dev)
                torch.cuda.synchronize()
                api.dev_to_mont(ctx, d, self.n, d)
                self.d_assign.append(d)
        # SNARK::encode does not depend on the witness: on one GPU a helper thread runs it on a second context (own stream and
        # scratch, this instance's priority) while this thread commits to the three assignments (ctypes releases the GIL; a Rust
        # shim would use std::thread::scope). VPIN_BENCH_OVERLAP_ENCODE=0 keeps the calls strictly
        # sequential. On a distributed context encode's commitments are sharded by the context's communicator, so it stays there.
        # Mode "1" is a measured alternative that lost: vPIN's my_lib_prove reads only encode's dense tables (the computation
        # commitment never enters the prover's transcript, VP/commit_test.rs:75), so the commitment's MSMs can run on a BACKGROUND
        # context (lowest stream priority, one MSM block per SM) underneath the proof - but 4.5 ms of MSM moved there slowed the
        # multiplier-bound round kernels of the proof by 7 ms (CNN A step 45.2 against 41.3 ms, profiles/r2_encode_overlap_modes.log).
        self.aux = self.bg = self.enc_pool = None
        self.enc_mode = os.environ.get("VPIN_BENCH_OVERLAP_ENCODE", "2")
        if getattr(ctx, "world", 1) == 1 and self.enc_mode in ("1", "2"):
            from concurrent.futures import ThreadPoolExecutor
            self.aux = api.Context(ctx.device, high_priority=high_priority)
            self.bg = api.Context(ctx.device, background=True) if self.enc_mode == "1" else None
            self.enc_pool = ThreadPoolExecutor(max_workers=1)
        L = self.gens.L
        self.d_pts = [torch.empty(32 * L, dtype=torch.uint8, device=dev) for _ in range(4)]
        self.d_blinds = [torch.empty(32 * L, dtype=torch.uint8, device=dev) for _ in range(3)]
        ctx.sync()

    def h2d_bytes(self):
        # COO triples + three assignments for the commitments + the assignment, commitment and blinds again for my_lib_prove
        L = self.gens.L
        return sum(a.nbytes for a in self.coo) + 4 * 32 * self.n + 2 * 32 * L * 2 + 64 * L + len(self.inputs)

    def encode(self, overlap=True, inst=None, gens=None):
        """SNARK::encode of the step -> (decommitment, commitment), each a callable that waits for its half: on the helper
        thread the tables are made first and the commitment right behind them (one worker), else both are ready at once"""
        from vpin_b200 import api
        inst, gens = inst or self.inst, gens or self.gens
        if overlap and self.aux is not None and self.enc_mode == "2":
            f = self.enc_pool.submit(api.SNARK.encode, inst, gens, self.aux)
            return (lambda: f.result()[1]), (lambda: f.result()[0])
        if overlap and self.aux is not None:
            f_tab = self.enc_pool.submit(api.encode_tables, inst, gens, self.aux)
            box = {}

            def get_decomm():  # called when the witness commitments are done: the commitment's MSMs start now, under the proof
                d = f_tab.result()
                box["f"] = self.enc_pool.submit(api.encode_commit, d, gens, self.bg)
                return d
            return get_decomm, (lambda: box["f"].result())
        comm, decomm = api.SNARK.encode(inst, gens)
        return (lambda: decomm), (lambda: comm)

    def close(self):
        if self.enc_pool is not None:
            self.enc_pool.shutdown()
        for c in (self.aux, self.bg):
            if c is not None:
                c.close()
        self.aux = self.bg = self.enc_pool = None

    def step_resident(self, ctx, seeds, overlap=True):
        from vpin_b200 import api
        sq, sp = seeds
        get_decomm, get_comm = self.encode(overlap)
        tape = api.RandomTape(b"\x02", sq)
        api.dev_poly_commit(ctx, self.gens, self.d_assign[0], self.n, tape, self.d_pts[0], self.d_blinds[0])
        api.dev_poly_commit(ctx, self.gens, self.d_assign[1], self.n, tape, self.d_pts[1], self.d_blinds[1])
        api.dev_poly_commit_with_blinds(ctx, self.gens, self.d_assign[2], self.n, self.d_blinds[0], self.d_blinds[1], self.d_pts[2],
                                        self.d_blinds[2])
        api.dev_commitments_add(ctx, self.d_pts[0], self.d_pts[1], self.gens.L, self.d_pts[3])
        decomm = get_decomm()
        wit = api.DeviceWitness(ctx, self.gens, self.d_assign[2], self.n, self.d_pts[3], self.d_blinds[2])
        proof = api.my_lib_prove_resident(self.inst, decomm, wit, self.inputs, self.gens, TRANSCRIPT_LABEL, sp)
        comm = get_comm()
        return comm, proof

    def step_e2e(self, ctx, seeds):
        """host buffers in, host bytes out — every call a Rust shim of the reference driver would make"""
        from vpin_b200 import api
        import numpy as np
        T = [time.time()]
        A, B, Cm = (np.frombuffer(t.numpy()[: n * api.COO_DTYPE.itemsize], dtype=api.COO_DTYPE) for t, n in self.h_coo)
        vp, vi, v = (t.numpy() for t in self.h_assign)
        sq, sp = seeds
        if self.aux is not None and self.enc_mode == "2":
            # Instance::new and SNARK::encode need no witness, the three commitments no instance: the helper thread builds and
            # encodes the instance on the second context while this thread commits to the assignments
            T.append(time.time())
            gens = api.SNARKGens(ctx, *self.dims); T.append(time.time())

            def build_and_encode():
                inst_ = api.Instance(self.aux, self.dims[0], self.dims[1], self.dims[2], A, B, Cm)
                comm_, decomm_ = api.SNARK.encode(inst_, gens, self.aux)
                return inst_, comm_, decomm_
            fut = self.enc_pool.submit(build_and_encode)
            get_decomm, get_comm = (lambda: fut.result()[2]), (lambda: fut.result()[1])
            inst = None
            T.append(time.time())
        else:
            inst = api.Instance(ctx, self.dims[0], self.dims[1], self.dims[2], A, B, Cm); T.append(time.time())
            gens = api.SNARKGens(ctx, *self.dims); T.append(time.time())
            get_decomm, get_comm = self.encode(True, inst, gens); T.append(time.time())
        tape = api.RandomTape(b"\x02", sq)
        c_para, b_para = api.dense_mlpoly_commit(ctx, gens, api._buf(vp), tape, n=self.n)
        c_input, b_input = api.dense_mlpoly_commit(ctx, gens, api._buf(vi), tape, n=self.n)
        c_vars, b_vars = api.my_dense_mlpoly_commit(ctx, gens, api._buf(v), b_para, b_input, n=self.n)
        combined = ctx.commitments_add(c_para, c_input)
        decomm = get_decomm(); T.append(time.time())
        if inst is None:
            inst = fut.result()[0]
        proof = api.my_lib_prove(inst, decomm, api._buf(v), self.inputs, gens, TRANSCRIPT_LABEL, combined, b_vars, sp, n=self.n, ctx=ctx)
        comm = get_comm(); T.append(time.time())
        fut = None
        del inst, decomm, gens, get_decomm, get_comm
        T.append(time.time())
        # per-call wall times of this step (Instance::new, SNARKGens::new, encode, commits, prove, handle release), kept for the
        # slowest step of the leg (bench line: e2e.slowest_step_calls_ms)
        self.last_e2e_calls_ms = [round(1e3 * (b - a), 2) for a, b in zip(T, T[1:])]
        return comm, proof, (c_para, c_input, c_vars)


# ------------------------------------------------------------------------------------------------------------ reference arm
def reference_build(wl):
    """the R1CS instances + assignments of the workload in the CPU port's format (built once, outside the timed steps: the
    metric starts where bench.py's own arm starts, after the R1CS exists)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    return {"add": O.build_point_add(*wl["add"]) if wl["add"] is not None else None, "mult": O.build_point_mult(*wl["mult"])}


def reference_step(wl, built, threads):
    """the whole workload on the CPU port (both instances at full size, one after the other like
    vPIN_proof_generation/src/main.rs:14-46); returns (seconds, per-instance seconds)"""
    import oracle_lib as O
    sq, sp = wl["seeds"]
    keys = ("gens", "SNARK::encode", "witness_commits", "SNARK::prove")
    total, parts = 0.0, {}
    if built["add"] is not None:
        f = O.Flow(built["add"], sq, sp, verify=False, threads=threads)
        parts["point_add"] = sum(f.times[k] for k in keys) / 1e3
        total += parts["point_add"]
    f = O.Flow(built["mult"], sq, sp, verify=False, threads=threads)
    parts["point_mult"] = sum(f.times[k] for k in keys) / 1e3
    parts["point_mult_phases_ms"] = {k: round(f.times[k], 1) for k in keys}
    total += parts["point_mult"]
    return total, parts


def reference_sample_text(wl, parts):
    return (f"full workload every step: point-add instance n={wl['n_add']} ({parts.get('point_add', 0.0):.2f} s) + point-mult instance "
            f"m={wl['m']} ({parts['point_mult']:.2f} s); MSM rows on all cores, everything else on one (as the reference: rayon only in commit_inner)")


def run_reference(args):
    rank, local_rank, world = rank_world()
    if rank != 0:
        return
    wl = make_workload(args.workload)
    threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    O.lib()
    built = reference_build(wl)
    t_start = time.time()
    warm_done, parts = 0, None
    first = None
    for _ in range(args.warmup):
        t0 = time.time()
        _, parts = reference_step(wl, built, threads)
        first = first or (time.time() - t0)
        warm_done += 1
        # the warm-up of a CPU arm has nothing to warm; stop it early when the budget would not hold the timed steps
        if (time.time() - t_start) + first * (args.steps + (args.warmup - warm_done)) > REF_BUDGET_S:
            break
    vals = []
    t0 = time.time()
    for k in range(args.steps):
        s_, parts = reference_step(wl, built, threads)
        vals.append(s_)
        per = (time.time() - t0) / len(vals)
        if k + 1 < args.steps and (time.time() - t_start) + per > REF_BUDGET_S:
            break
    wall = time.time() - t0
    value = sum(vals) / len(vals)
    sample = reference_sample_text(wl, parts)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": warm_done, "ms_per_step": 1e3 * wall / len(vals), "higher_is_better": False, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "u64 (4x64 Montgomery F_l, 5x51 F_p)", "data": "synthetic",
        "config": {"workload": f"cnn{wl['tag']}" if len(wl['tag']) == 1 else wl["tag"], "point_mults": wl["m"], "point_adds": wl["n_add"],
                   "sample": sample},
        "requested": {"steps": args.steps, "warmup": args.warmup, "budget_s": REF_BUDGET_S},
        "per_instance_s": {k: v for k, v in parts.items() if not isinstance(v, dict)},
        "phases_ms_point_mult": parts["point_mult_phases_ms"],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------ B200 arm
def load_ncu_facts():
    """per-launch facts of the dominant kernel that only a profiler can give (DRAM bytes, multiply-pipe utilisation), taken
    from the committed `ncu --set full` capture of the same kernel (profiles/); None when the file is absent"""
    p = os.path.join(ROOT, "profiles", "r2_ncu_msm_accumulate_facts.json")
    try:
        return json.load(open(p))
    except (OSError, ValueError):
        return None


class Leg:
    """One measured arrangement of the workload on this rank: the network's R1CS instances (point additions, point
    multiplications) are independent proofs (vPIN_proof_generation/src/main.rs:14-46 runs them one after the other):
    each gets its own context — stream, generator tables, NCCL communicator when `distributed` — and its own host thread,
    so the latency-bound rounds of one overlap the other's kernels."""

    def __init__(self, args, torch, dist, wl, distributed, device_built=False):
        from concurrent.futures import ThreadPoolExecutor
        from vpin_b200 import api
        wls = wl if isinstance(wl, list) else [wl]  # several networks: all of them are proved side by side in one step
        self.args, self.torch, self.dist, self.wl = args, torch, dist, wls[0]
        self.rank, self.local_rank, self.world = rank_world()
        self.dev = torch.device("cuda", self.local_rank)
        prio = os.environ.get("VPIN_BENCH_PRIORITY", "1") == "1"
        self.states = []
        for w in wls:
            # point-mult first: its generator tables (shared per device and label) then also serve the point-add instance
            builders = [("point_mult", (lambda c, w=w: api.point_mult_device(c, *w["mult"])) if device_built else
                         (lambda c, w=w: api.point_mult(c, *w["mult"])))] + \
                       ([("point_add", lambda c, w=w: api.point_addition(c, *w["add"]))] if w["add"] is not None else [])
            for kind, build in builders:
                # the point-mult proof is the critical path of the step: its stream gets the urgent priority, so the point-add
                # instance's kernels fill the gaps instead of delaying its latency-bound rounds
                c = api.Context(self.local_rank, high_priority=prio and kind == "point_mult")
                # ONE communicator per process: the point-mult proof (the critical path, > 95 % of the work) is the one that is
                # sharded; the small point-add proof runs beside it on every rank without a communicator. Two NCCL communicators
                # driven concurrently from two host threads may issue their collectives in different orders on different ranks,
                # which NCCL does not tolerate (observed: a hang at N = 2).
                if distributed and kind == "point_mult":
                    c.init_distributed(self.rank, self.world, dist)
                self.states.append(InstanceState(c, kind, build(c), torch, device_built=device_built and kind == "point_mult",
                                                 high_priority=prio and kind == "point_mult"))
        self.states.sort(key=lambda s_: s_.kind)  # point_add before point_mult (the order the JSON line lists them in)
        self.ctx = self.states[-1].ctx  # a point-mult context
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.pool = ThreadPoolExecutor(max_workers=len(self.states))

    def sync_all(self):
        self.torch.cuda.synchronize()
        for s_ in self.states:
            s_.ctx.sync()

    def barrier(self):
        self.sync_all()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run_all(self, fn):
        futs = [self.pool.submit(fn, s_) for s_ in self.states]
        return [f.result() for f in futs]

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def one_step_resident(self):
        self.flush.zero_()  # L2 flush between steps (and the working set, ~2 GB, is far larger than L2 anyway)
        self.torch.cuda.synchronize()
        seeds = self.wl["seeds"]
        return self.run_all(lambda s_: s_.step_resident(s_.ctx, seeds))

    def one_step_e2e(self):
        self.flush.zero_()
        self.torch.cuda.synchronize()
        seeds = self.wl["seeds"]
        return self.run_all(lambda s_: s_.step_e2e(s_.ctx, seeds))

    def time_resident(self, sample_clocks):
        """W untimed steps, then exactly K steps between barriers, CUDA events on the library's stream, max over ranks"""
        torch, args = self.torch, self.args
        for _ in range(args.warmup):
            self.first = self.one_step_resident()
        self.barrier()
        clocks = ClockSampler(self.local_rank) if (sample_clocks and self.rank == 0) else None
        import gc
        gc.collect()
        gc.disable()  # the harness is Python: keep its cyclic collector (7 ms pauses) out of the timed steps
        launches = lambda: sum(s_.ctx.kernel_launches + sum(c.kernel_launches for c in (s_.aux, s_.bg) if c is not None) for s_ in self.states)
        l0 = launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(self.stream)  # the device is idle here (barrier) ...
        for _ in range(args.steps):
            self.last = self.one_step_resident()
        self.sync_all()         # ... and here: every context's stream has drained before the closing event is recorded
        e1.record(self.stream)
        self.barrier()
        wall = time.time() - t0
        gc.enable()
        dev_ms = e0.elapsed_time(e1)
        out = {"launches": launches() - l0, "phases": self.ctx.phase_times(),
               "clocks": clocks.stop() if clocks else None,
               "step_s": self.max_over_ranks(dev_ms / 1e3 / args.steps), "wall_step_s": self.max_over_ranks(wall / args.steps)}
        assert [p for _, p in self.last] == [p for _, p in self.first], "proof bytes changed between steps (must be deterministic)"
        return out

    def profile_pass(self):
        """per-kernel-class device times: the same K steps again with a CUDA-event scope around every kernel class, the
        instances one after the other so that the scopes time one kernel at a time. Kept out of the timed region: the
        event records between the two contexts' launches slow a concurrent step down by 2x."""
        torch, args = self.torch, self.args
        prof, madds = {}, 0
        for s_ in self.states:
            s_.ctx.profile_enable(True, 32768.0)
        self.barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(self.stream)
        seeds = self.wl["seeds"]
        for _ in range(args.steps):
            self.flush.zero_()
            torch.cuda.synchronize()
            prof_out = [s_.step_resident(s_.ctx, seeds, overlap=False) for s_ in self.states]
        self.sync_all()
        p1.record(self.stream)
        self.barrier()
        prof_ms = p0.elapsed_time(p1)
        assert [p for _, p in prof_out] == [p for _, p in self.first]
        for s_ in self.states:
            p_, m_ = s_.ctx.profile_read()
            madds += m_
            for k_, v_ in p_.items():
                a_ = prof.setdefault(k_, {key: 0 for key in v_})
                for key in v_:
                    a_[key] += v_[key]
            s_.ctx.profile_enable(False)
        return prof, madds, prof_ms

    def time_e2e(self):
        """the same step through the host-buffer C ABI: pinned host inputs up, proof bytes down, every step"""
        args = self.args
        for _ in range(max(1, args.warmup)):
            e2e_out = self.one_step_e2e()
        self.barrier()
        import gc
        gc.collect()
        gc.disable()
        per_step = []
        t0 = time.time()
        slowest, slowest_calls = 0.0, None
        for _ in range(args.steps):
            t1 = time.time()
            e2e_out = self.one_step_e2e()
            per_step.append(time.time() - t1)
            if per_step[-1] > slowest:
                slowest, slowest_calls = per_step[-1], {s_.kind: s_.last_e2e_calls_ms for s_ in self.states}
        self.barrier()
        e2e_s = self.max_over_ranks((time.time() - t0) / args.steps)
        gc.enable()
        self.e2e_steps_ms = [round(1e3 * x, 2) for x in per_step]
        self.e2e_slowest_calls = slowest_calls
        for (c1, p1), (c2, p2, _) in zip(self.last, e2e_out):
            assert c1 == c2 and p1 == p2, "resident and host-buffer legs disagree"
        h2d = sum(s.h2d_bytes() for s in self.states)
        d2h = sum(len(c) + len(p) + 3 * 32 * s.gens.L + 3 * 32 * s.gens.L for (c, p, _), s in zip(e2e_out, self.states))
        return e2e_s, h2d, d2h

    def set_shard_sumcheck(self, on):
        for s_ in self.states:
            s_.ctx.set_shard_sumcheck(on)

    def parity(self, golden):
        """sha256 of what the last timed step produced (computation commitment + proof per instance) against the committed
        digests of the CPU oracle for this shape (tests/golden/golden_named.json), and - on several ranks - identical on all"""
        import hashlib
        out = {"instances": [], "matches_golden": True, "identical_on_all_ranks": True}
        mine = []
        for s_, (comm, proof) in zip(self.states, self.last):
            hp, hc = hashlib.sha256(proof).hexdigest(), hashlib.sha256(comm).hexdigest()
            mine.append(hp + hc)
            want = golden.get((self.wl["tag"], s_.kind))
            ok = bool(want) and want["proof_sha256"] == hp and want["comm_sha256"] == hc and want["proof_len"] == len(proof)
            out["instances"].append({"kind": s_.kind, "proof_bytes": len(proof), "proof_sha256": hp[:16], "golden": "match" if ok else
                                     ("none committed for this shape" if not want else "MISMATCH")})
            if want and not ok:
                out["matches_golden"] = False
            if not want:
                out["matches_golden"] = None if out["matches_golden"] is True else out["matches_golden"]
        if self.world > 1:
            box = [None] * self.world
            self.dist.all_gather_object(box, mine)
            out["identical_on_all_ranks"] = all(b == box[0] for b in box)
        return out

    def close(self):
        # handles (gens, instances, decommitments) must go before the context that owns their stream
        import gc
        ctxs = [s_.ctx for s_ in self.states]
        self.pool.shutdown()
        for s_ in self.states:
            s_.close()
        self.states = self.first = self.last = None
        self.ctx = None
        gc.collect()
        for c in ctxs:
            c.close()


def load_golden():
    """(tag, kind) -> digests of the CPU oracle's flow for the named shapes (tests/golden/make_golden.py named)"""
    try:
        d = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_named.json")))
    except (OSError, ValueError):
        return {}
    return {(c["tag"], c["kind"]): c for c in d["cases"]}


def class_rooflines(prof, madds, prof_ms, steps, imad_peak, hbm_peak):
    """per-kernel-class roofline entries of a profile pass (Leg.profile_pass): the MSM against the live IMAD.WIDE peak (executed
    additions x 504), everything else its algorithmic bytes against the measured HBM copy rate"""
    out = []
    for name, p in prof.items():
        if p["ms"] <= 0:
            continue
        ent = {"kernel": name, "launches": p["launches"] // steps, "ms_per_step": p["ms"] / steps, "share_of_step": p["ms"] / prof_ms}
        if name == "msm_accumulate":
            ent.update(bound="imad", achieved=madds * 504.0 / (p["ms"] * 1e-3) / 1e12, peak=imad_peak / 1e12, unit="TMAC/s")
        elif p["bytes"] > 0:
            ent.update(bound="hbm", achieved=p["bytes"] / (p["ms"] * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s")
        else:
            continue
        ent["frac"] = ent["achieved"] / ent["peak"]
        out.append(ent)
    out.sort(key=lambda e: -e["ms_per_step"])
    return out


def short_leg(args, torch, dist, tag, distributed, golden, steps=10, warmup=3, rooflines_with=None):
    """another named shape, same measurement (steps / warm-up reduced): seconds per step + parity against the golden digests"""
    import copy
    a2 = copy.copy(args)
    a2.steps, a2.warmup = min(args.steps, steps), min(args.warmup, warmup)
    # (the point-mult instance is expanded on the device: a short leg times the resident step only and needs no host copy of the
    # 78 M COO triples of a LeNet-layer-5 instance)
    leg = Leg(a2, torch, dist, make_workload(tag), distributed=distributed, device_built=True)
    r = leg.time_resident(sample_clocks=False)
    par = leg.parity(golden)
    ph = r["phases"]
    out = {"value": r["step_s"], "unit": UNIT, "steps": a2.steps, "warmup": a2.warmup,
           "instances": [{"kind": s_.kind, "num_cons": s_.dims[0], "nnz_param": s_.dims[3], "hyrax_grid": [s_.gens.L, s_.gens.R]} for s_ in leg.states],
           "matches_golden": par["matches_golden"], "identical_on_all_ranks": par["identical_on_all_ranks"],
           "snark_prove_ms_point_mult": ph.get("SNARK::prove")}
    if rooflines_with and not distributed:
        # the per-kernel-class rooflines at a size where the F_l kernels stream instead of waiting for launches (CNN E: 2^22
        # constraints) - same scopes, same byte counts as the headline's `rooflines`
        try:
            prof, madds, prof_ms = leg.profile_pass()
            out["rooflines"] = class_rooflines(prof, madds, prof_ms, a2.steps, *rooflines_with)[:8]
        except Exception as e:  # noqa: BLE001
            out["rooflines"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    leg.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 prover has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hbm_peak, peak_src = load_peaks()
    golden = load_golden()
    # host threads per rank: 2 instance threads (they spin on the round slots), 2 delta workers, 2 x 2 spinning helpers for
    # the sigma-protocol commitments; without ~10 cores per rank the helpers would only steal from the spinning main threads
    cores_per_rank = (os.cpu_count() or 1) // world
    if cores_per_rank < 10:
        os.environ.setdefault("VPIN_HOST_HELPERS", "1" if cores_per_rank >= 8 else "0")

    # ---- headline arrangement: ONE proof of the named network (both instances). N = 1: on one B200. N > 1: the same proof
    # with its work sharded over the N GPUs (strong scaling: total work fixed) - every rank runs the same calls on the same
    # inputs, the Hyrax commitment rows are split across the ranks and exchanged by NCCL all-gathers, the transcript is
    # replayed identically everywhere. The proof bytes must equal the single-GPU ones: checked against the committed golden
    # digests of the CPU oracle and across the ranks.
    wl = make_workload(args.workload)
    leg = Leg(args, torch, dist, wl, distributed=world > 1)
    gens_cold_s = max(s_.gens_s for s_ in leg.states)  # SNARKGens::new with nothing cached: SHAKE stream, hash-to-group, fixed-base tables
    imad_forms = leg.ctx.imad_peak_forms()
    imad_peak = max(imad_forms)
    res = leg.time_resident(sample_clocks=True)
    step_s = res["step_s"]
    parity = leg.parity(golden)
    prof, madds, prof_ms = ({}, 0, 0.0) if args.no_profile else leg.profile_pass()
    e2e_s, h2d, d2h = leg.time_e2e()
    e2e_steps_ms = leg.e2e_steps_ms
    e2e_slowest_calls = leg.e2e_slowest_calls
    msm = msm_uniform_bench(leg.ctx, torch, leg.dev, leg.stream, imad_peak, world)
    # witness expansion + R1CS emission of the point-mult instance on the device (vpin_build_point_mult_device), assignments left
    # in HBM: the step of vPIN's timed region that precedes the prover (proof_point_mult.rs:24, point_mult.rs:7-664)
    from vpin_b200 import api as _api
    t_build = []
    for _ in range(3):
        leg.sync_all()
        t0 = time.time()
        built = _api.point_mult_device(leg.ctx, *wl["mult"])
        leg.ctx.sync()
        t_build.append(time.time() - t0)
        del built
    witness_build = {"point_mult_build_s": min(t_build), "what": "witness expansion (256 inversions per multiplication) + COO emission + "
                     "Instance::new on the device from the JSON-level inputs (weights, point coordinates); not part of `value`"}
    instances = [{"kind": s.kind, "num_cons": s.dims[0], "num_vars": s.dims[1], "nnz_param": s.dims[3], "hyrax_grid": [s.gens.L, s.gens.R]}
                 for s in leg.states]
    # ---- N > 1, beside the headline: the same proof with ONLY the commitment rows sharded (vpin_ctx_set_shard_sumcheck(ctx, 0):
    # hash layers, product trees and every sumcheck round replicated on all ranks). Same bytes.
    rows_only = None
    if world > 1:
        leg.set_shard_sumcheck(0)
        r2 = leg.time_resident(sample_clocks=False)
        p2 = leg.parity(golden)
        rows_only = {"value": r2["step_s"], "unit": UNIT, "matches_golden": p2["matches_golden"],
                     "identical_on_all_ranks": p2["identical_on_all_ranks"], "phases_ms_point_mult": r2["phases"],
                     "what": "commitment rows sharded, everything else replicated (the round-1 arrangement)"}
        leg.set_shard_sumcheck(-1)
    leg.close()

    # ---- serving mode on one GPU: B different networks proved side by side. A single CNN-A proof is latency-bound (about half
    # of its 50 ms the GPU waits for the host's next Fiat-Shamir challenge), so concurrent proofs fill each other's gaps; the
    # generator tables are shared by all contexts of the process. Reported beside the headline, not instead of it.
    concurrent = None
    B = env_int("VPIN_BENCH_CONCURRENT", 4)
    if world == 1 and B > 1:
        saved = os.environ.get("VPIN_HOST_HELPERS")
        os.environ["VPIN_HOST_HELPERS"] = "0"  # 2 B proving threads + their delta workers already occupy the cores
        try:  # an auxiliary leg must never cost the headline line
            legc = Leg(args, torch, dist, [make_workload(args.workload, replica=r) for r in range(B)], distributed=False)
            rc = legc.time_resident(sample_clocks=False)
            concurrent = {"networks": B, "s_per_step": rc["step_s"], "networks_per_s": B / rc["step_s"],
                          "what": f"{B} different {args.workload} networks ({2 * B} proofs) in flight on one B200, one context and host thread per proof"}
            legc.close()
        except Exception as e:  # noqa: BLE001
            concurrent = {"error": f"{type(e).__name__}: {e}"[:300]}
        if saved is None:
            os.environ.pop("VPIN_HOST_HELPERS", None)
        else:
            os.environ["VPIN_HOST_HELPERS"] = saved

    # ---- N > 1, beside the headline: N replicas. vPIN proves independent networks (one per inference), so the arrangement
    # with no data-path collective at all is one network per rank (different witnesses per rank), weak scaling.
    replicas = None
    if world > 1:
        try:
            legr = Leg(args, torch, dist, make_workload(args.workload, replica=rank), distributed=False)
            rr = legr.time_resident(sample_clocks=False)
            replicas = {"value": rr["step_s"], "unit": UNIT, "scaling": "weak", "networks_per_step": world, "networks_per_s": world / rr["step_s"],
                        "what": f"{world} different networks of the shape, one per rank, no data-path collective; seconds until the slowest rank is done"}
            legr.close()
        except Exception as e:  # noqa: BLE001
            replicas = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- the other named shapes of BASELINE.json (conv 3/5/7 sweep, CNN E), same arrangement as the headline, fewer steps
    other = {}
    # (LeNet layer 5, the 230 GB / 2^25-constraint instance, joins at 8 GPUs - BASELINE.json's fifth config; VPIN_BENCH_OTHER overrides)
    extra = os.environ.get("VPIN_BENCH_OTHER", "conv3,conv5,conv7,E" + (",L5" if world >= 8 else ""))
    for tag in [t for t in extra.split(",") if t and t != args.workload]:
        try:
            other[tag] = short_leg(args, torch, dist, tag, world > 1, golden,
                                   rooflines_with=(imad_peak, hbm_peak) if tag == "E" and not args.no_profile else None)
        except Exception as e:  # noqa: BLE001  (an auxiliary leg must never cost the headline line)
            other[tag] = {"error": f"{type(e).__name__}: {e}"[:300]}
            if world > 1:
                break  # the ranks may no longer be in step

    # ---- roofline of the dominant kernel class -----------------------------------------------------------------------
    facts = load_ncu_facts()
    rooflines = []
    for name, p in prof.items():
        if p["ms"] <= 0:
            continue
        ent = {"kernel": name, "launches": p["launches"] // args.steps, "ms_per_step": p["ms"] / args.steps,
               "share_of_step": p["ms"] / prof_ms}
        ent["traffic"] = None
        if name == "msm_accumulate":
            macs = madds * 504.0  # 7 F_p multiplications of 72 multiply-accumulates per mixed addition
            ent.update(bound="imad", achieved=macs / (p["ms"] * 1e-3) / 1e12, peak=imad_peak / 1e12, unit="TMAC/s",
                       madds_per_step=madds // args.steps, points_per_step=p["units"] / args.steps)
            if facts:
                if wl["tag"] == "A" and world == 1 and facts.get("step_launches") == p["launches"] // args.steps:
                    # measured: ncu DRAM bytes of the k_msm_accumulate launches of one CNN-A bench step, averaged per launch
                    ent["traffic"] = facts["step_dram_bytes_per_launch_avg"]
                    ent["traffic_how"] = "dram__bytes_read.sum + dram__bytes_write.sum of the %d launches of one step / %d (profiles/r2_msm_step_traffic.csv)" % (
                        facts["step_launches"], facts["step_launches"])
                else:
                    # DRAM bytes per mixed addition from the ncu --set full capture x the additions one launch executes here
                    ent["traffic"] = facts["dram_bytes_per_madd"] * (madds / max(1, p["launches"]))
                    ent["traffic_how"] = "bytes per mixed addition of the ncu --set full capture x additions per launch here"
                ent["algorithmic_bytes_per_launch"] = 98.0 * madds / max(1, p["launches"])  # 96-byte table entry + 2-byte digit per addition
                ent["traffic_source"] = facts["source"]
                ent["multiply_pipe_util_ncu"] = facts["fmaheavy_pipe_util"]
        elif p["bytes"] > 0:
            ent.update(bound="hbm", achieved=p["bytes"] / (p["ms"] * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s")
        else:
            continue
        ent["frac"] = ent["achieved"] / ent["peak"]
        rooflines.append(ent)
    rooflines.sort(key=lambda e: -e["ms_per_step"])
    top = dict(rooflines[0]) if rooflines else None
    if top:
        top["peak_source"] = peak_src if top["bound"] == "hbm" else \
            ("measured live: dependency-free stream of real IMAD.WIDE.U32 (vpin_imad_peak_forms: %.2f T/s plain product + ALU combine, "
             "%.2f T/s single-instruction multiply-accumulate; the better one). IMAD.WIDE is a half-rate instruction on sm_100a; the "
             "round-1 peak of 18.4 T/s timed IADD3 pairs (its product had been hoisted out of the loop)" % (imad_forms[0] / 1e12, imad_forms[1] / 1e12))

    tag = f"cnn{wl['tag']}" if len(wl['tag']) == 1 else wl["tag"]
    line = {
        "metric": METRIC, "value": step_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * step_s, "higher_is_better": False, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "u32 limbs (8x32 Montgomery F_l, 8x32 F_p; IMAD.WIDE)", "data": "synthetic",
        "config": {"workload": tag, "point_mults": wl["m"], "point_adds": wl["n_add"], "instances": instances,
                   "l2": "256 MB buffer written between steps; working set ~2 GB >> 126 MB L2",
                   "concurrency": "the network's independent instances are proved concurrently (one context + host thread each); on one GPU "
                                  "SNARK::encode runs on a second context of its instance beside the three witness commitments",
                   "parallelism": "1 GPU" if world == 1 else
                                  f"ONE {tag} proof on {world} GPUs: every rank replays the transcript; the Hyrax commitment rows are split "
                                  "across the ranks (NCCL all-gather of 32 B per row); the product circuits of the memory check are dealt "
                                  "to the ranks (hash layers, trees, the sumcheck rounds of the large layers: one <= 3 KB all-gather per "
                                  "round) and the dense evaluations are computed in slices; layers below 2^17 items stay replicated"},
        "host": {"cores": os.cpu_count(), "helpers_per_prover": int(os.environ.get("VPIN_HOST_HELPERS", "2"))},
        "parity": parity,
        "sharded_equals_unsharded": (parity["matches_golden"] is True and parity["identical_on_all_ranks"]) if world > 1 else None,
        "wall_s_per_step": res["wall_step_s"],
        "gpu_launches": res["launches"],
        "clocks": res["clocks"],
        "e2e": {"value": e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps_ms": e2e_steps_ms,
                "slowest_step_calls_ms": e2e_slowest_calls,
                "calls": ["Instance::new (on one GPU: inside the helper thread, 0 here)", "SNARKGens::new",
                          "Instance::new + SNARK::encode submitted to the second context",
                          "3 commits + combine, then the join with the helper thread", "my_lib_prove", "release"]},
        "roofline": top,
        "roofline_pass": {"ms_per_step": prof_ms / args.steps if prof_ms else None,
                          "how": "same K steps repeated after the timed region with CUDA-event scopes on the launching stream, instances sequential"},
        "rooflines": rooflines[:8],
        "imad_peak_forms_tmacs": {"plain_product_plus_alu_combine": imad_forms[0] / 1e12, "single_instruction_mac": imad_forms[1] / 1e12},
        "msm": msm,
        "one_proof_rows_only": rows_only,
        "replicas": replicas,
        "concurrent_proofs": concurrent,
        "other_configs": other,
        "witness_build": witness_build,
        "gens_cold_s": gens_cold_s,
        "phases_ms_point_mult": res["phases"],
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        built = reference_build(wl)
        s, parts = reference_step(wl, built, threads)
        line["cpu_baseline"] = {"value": s, "unit": UNIT, "cores": threads, "kind": "port", "sample": reference_sample_text(wl, parts)}
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def msm_uniform_bench(ctx, torch, dev, stream, imad_peak, world=1):
    """Hyrax commitment of 2^22 uniform full-width scalars (2048 rows x 2048 generators), device resident; on a distributed
    context the rows are sharded over the ranks (the fraction of the integer roofline is per GPU)."""
    from vpin_b200 import api
    ell = 22
    n = 1 << ell
    g = torch.Generator(device="cpu").manual_seed(0x7650494E)
    z = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g)
    z[:, 31] &= 0x0F  # < 2^252 < l: canonical
    d = z.to(dev)
    torch.cuda.synchronize()
    api.dev_to_mont(ctx, d, n, d)
    out = torch.empty(32 * (1 << (ell // 2)), dtype=torch.uint8, device=dev)
    label = b"gens_r1cs_eval"
    lib = api.lib()
    for _ in range(2):
        ctx.check(lib.vpin_dev_hyrax_commit(ctx._h, label, C.c_void_p(d.data_ptr()), C.c_uint64(n), None, C.c_void_p(out.data_ptr())))
    ctx.sync()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ctx.check(lib.vpin_dev_hyrax_commit(ctx._h, label, C.c_void_p(d.data_ptr()), C.c_uint64(n), None, C.c_void_p(out.data_ptr())))
    e1.record(stream)
    ctx.sync()
    sec = e0.elapsed_time(e1) / 1e3 / reps
    return {"workload": "Hyrax commit, 2^22 uniform full-width scalars, 2048x2048", "mpoints_per_s": n / sec / 1e6,
            "algorithmic_macs_per_point": 8064, "algorithmic_frac_of_imad_peak": n * 8064 / sec / imad_peak / world,
            "imad_peak_tmacs": imad_peak / 1e12, "gpus": world}


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line, on the real stdout (fd 1 is pointed at stderr while the run is in progress so that library
    banners — e.g. NCCL's version line — cannot get in front of it)"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="A")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel-class CUDA-event scopes (rooflines come out empty)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
