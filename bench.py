#!/usr/bin/env python3
"""bench.py -- headline benchmark of the gate-application hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W          # our CUDA engine
    python bench.py --impl reference [...]                 # the reference's own CPU code

One "step" = one pass of the hot path over one synthetic input: the reference's
qc_quantum_fourier_transform() on a resident state (30 qubits at N=1: 465 gates,
16 GiB of amplitudes; 30+log2(N) qubits sharded over N GPUs, 16 GiB per GPU).
`value` is whole-job throughput in gates/s, normalised to 2^30-amplitude shards:
(gates in the circuit) x (shards) / time.  At N=1 that is plain gates/s on the
30-qubit QFT, the figure BASELINE.json's metric names.

Keys beyond the base contract: `roofline` (dominant kernel = the fused tile pass,
HBM-bound; algorithmic bytes 32*2^nl per launch, duration from CUDA events inside
the engine), `cpu_baseline` (the unmodified reference, oracle/_ref, on this box's
host cores, bounded sample), `e2e` (create -> QFT -> read -> destroy through the
public C API of include/qcs.h), `clocks`, `gpu_launches`.

Nothing here reads /root/reference; the reference arm runs the prebuilt
oracle/_ref/*.so (built in the CPU container by `make -C oracle ref`).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = sys.stdout
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def qft_gate_count(n: int) -> int:
    return n * (n + 1) // 2


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = []
        smax = None
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU reference
def _mem_available_gib() -> float:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 2 ** 20
    except OSError:
        pass
    return 0.0


def _cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_sample(n_target: int, budget_s: float = 20.0) -> dict:
    """Times the first gates of the n_target-qubit QFT on the UNMODIFIED reference
    (oracle/_ref), sequential mode on one core and OpenMP mode on all cores."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    if not po.ref_available("seq"):
        # the reference library was not built (no /root/reference at build time): time our C port
        orc = po.Oracle(24, "reference")
        t0 = time.perf_counter()
        g = 0
        for q in range(6):
            orc.h(q); g += 1
        dt = time.perf_counter() - t0
        orc.close()
        gps = g / dt / 2 ** (n_target - 24)
        return {"value": gps, "unit": "gates/s", "cores": 1, "kind": "port",
                "sample": f"6 H gates at 24 qubits on oracle/qcs_oracle.c, scaled by 2^-{n_target - 24}"}
    # the reference holds two buffers: 32 * 2^n bytes
    n = n_target
    while n > 20 and 32 * 2 ** n / 2 ** 30 > 0.45 * _mem_available_gib():
        n -= 1
    scale = 2.0 ** (n - n_target)  # per-amplitude gate cost is flat once the state exceeds the caches
    per_gate_est = 6.0 * 2.0 ** (n - 30)  # s, survey-time single-core figure
    n_gates = max(3, min(24, int(budget_s / max(per_gate_est, 1e-3))))
    out = {}
    # the reference's sequential, QCS_SIMD_ONLY, QCS_CPU_OPENMP and QCS_MULTI_THREAD (pthread pool,
    # hard-capped at 4 threads: reference src/qcs.c:38) builds
    # ... and "seqlib": sequential mode built the way the reference's own Makefile builds its library
    # (one object per source file, c_mul / c_add not inlined) -- reported, ~10x slower than the
    # single-translation-unit builds above, which are how the reference's README numbers were made
    for mode in ("seq", "simd", "omp", "mt", "seqlib"):
        if not po.ref_available(mode):
            continue
        n_gates_mode, n_mode = n_gates, n
        if mode == "seqlib":  # ~35 s per gate at 30 qubits: sampled on a 26-qubit state (1 GiB, far beyond the caches)
            n_gates_mode, n_mode = 3, min(n, 26)
        if mode == "omp":
            os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        ref = po.RefLib(n_mode, mode)
        # warm the pages of both buffers with one untimed gate
        ref.h(n_mode - 1)
        done = 0
        t0 = time.perf_counter()
        i, j = 0, 1
        ref.h(0); done += 1
        while done < n_gates_mode:
            ref.cphase(j, i, math.pi / float(1 << (j - i))); done += 1
            j += 1
            if j >= n_mode:
                i += 1; j = i + 1
                ref.h(i); done += 1
        dt = time.perf_counter() - t0
        ref.close()
        out[mode] = {"gates": done, "seconds": dt, "sample_qubits": n_mode, "gates_per_s_at_sample_width": done / dt,
                     "gates_per_s": done / dt * 2.0 ** (n_mode - n_target),
                     "threads": cores if mode == "omp" else 4 if mode == "mt" else 1,
                     "build": "multi-TU library (reference Makefile layout)" if mode == "seqlib"
                              else "single translation unit (the reference's single-header usage)"}
    best = max((m for m in out if m != "seqlib"), key=lambda m: out[m]["gates_per_s"])
    return {"value": out[best]["gates_per_s"], "unit": "gates/s",
            "cores": out[best]["threads"], "kind": "reference",
            "sample": (f"first {out[best]['gates']} gates of the {n}-qubit QFT through the reference's own "
                       f"qc_h/qc_cphase (oracle/_ref, mode {best}"
                       + (f"; scaled by 2^{n - n_target} to {n_target} qubits" if n != n_target else "")
                       + "), qc_create not timed"),
            "modes": out, "host_cores": cores, "cpu_model": _cpu_model()}


class RefRunner:
    """The unmodified reference (oracle/_ref) walking through the n-qubit QFT gate by gate."""

    def __init__(self, n: int, mode: str):
        from oracle import pyoracle as po
        self.n, self.mode = n, mode
        self.ref = po.RefLib(n, mode)
        self.ref.h(n - 1)          # touch both buffers once (page faults are not gate time)
        self.i, self.j = 0, 0      # next gate: H(i) if j == 0 else CPHASE(i + j -> i)

    def step(self, gates: int) -> float:
        t0 = time.perf_counter()
        for _ in range(gates):
            if self.j == 0:
                self.ref.h(self.i)
            else:
                self.ref.cphase(self.i + self.j, self.i, math.pi / float(1 << self.j))
            self.j += 1
            if self.i + self.j >= self.n:
                self.i, self.j = (self.i + 1) % self.n, 0
        return time.perf_counter() - t0

    def close(self):
        self.ref.close()


def reference_arm(args, rank: int, world: int) -> None:
    """`--impl reference`: the reference's own CPU implementation of the same workload on this
    box's host cores; each step is a bounded sample (a few gates of the QFT)."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    n_target = args.qubits or (30 + int(math.log2(world)))
    cores = os.cpu_count() or 1
    if not po.ref_available("seq"):
        base = run_reference_sample(n_target, 10.0)
        value, sample, kind, used, modes = base["value"], base["sample"], base["kind"], base["cores"], None
    else:
        n = n_target
        while n > 20 and 32 * 2 ** n / 2 ** 30 > 0.45 * _mem_available_gib():
            n -= 1
        scale = 2.0 ** (n - n_target)
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        # Which of the reference's CPU modes is fastest on this host is decided on a 26-qubit state
        # (1 GiB: far beyond the caches, so the per-amplitude cost is the full-width one) -- keeping
        # four 32 GiB states alive to ask the same question would not fit every box.  Modes:
        # sequential, QCS_SIMD_ONLY, QCS_CPU_OPENMP (all cores), QCS_MULTI_THREAD (pthread pool, 4
        # threads: the reference's own cap, src/qcs.c:38).
        n_probe = min(n, 26)
        probe = {}
        for m in ("seq", "simd", "omp", "mt"):
            if not po.ref_available(m):
                continue
            r = RefRunner(n_probe, m)
            r.step(1)
            probe[m] = r.step(2) / 2.0 * 2.0 ** (n - n_probe)   # seconds per gate at width n
            r.close()
        best = min(probe, key=probe.get)
        # Every step times AT LEAST 8 consecutive gates (an H and the controlled phases behind it).  At
        # the full width one gate costs ~2 s on this class of host, so 8 gates x (steps + warm-up)
        # would not end "within a few minutes": the sample is taken on the widest state for which it
        # does (never below 26 qubits = 1 GiB, far beyond the caches: the per-amplitude cost of a gate
        # is the full-width one) and scaled by the ratio of the state sizes.
        total_steps = max(1, args.steps + args.warmup)
        g = 8
        n_run = n
        while n_run > 26 and g * total_steps * probe[best] * 2.0 ** (n_run - n) > 100.0:
            n_run -= 1
        scale = 2.0 ** (n_run - n_target)
        runner = RefRunner(n_run, best)
        times = [runner.step(g) for _ in range(total_steps)][args.warmup:]
        runner.close()
        value = g * len(times) / sum(times) * scale
        used = {"omp": cores, "mt": min(4, cores)}.get(best, 1)
        kind = "reference"
        modes = {m: {"seconds_per_gate_probe": t} for m, t in probe.items()}
        sample = (f"{g} consecutive gates of the {n_run}-qubit QFT per step through the reference's own "
                  f"qc_h/qc_cphase (oracle/_ref mode {best}, {used} thread(s) of {cores} cores"
                  + (f"; scaled by 2^{n_run - n_target} to {n_target} qubits" if n_run != n_target else "")
                  + "); qc_create not timed")
    # same unit as our arm: gates x 2^30-amplitude shards per second (at N = 1 plain gates/s of the
    # 30-qubit QFT; at N > 1 the reference sweeps N shards' worth of amplitudes per gate)
    value *= world
    line = {
        "impl": "reference", "metric": "gates/s", "value": value, "unit": "gates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * qft_gate_count(n_target) * world / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n_target}-qubit QFT (qc_quantum_fourier_transform), "
                               f"{qft_gate_count(n_target)} gates; reference CPU implementation, "
                               "bounded sample per step",
                   "value_definition": "gates x shards / time, shards of 2^30 amplitudes (our arm's definition)"},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "modes": modes,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ----------------------------------------------------------------------------- sharded parity
# The driver's GPU test box has one GPU, so tests/test_dist_gpu.py is skipped there.  A multi-rank
# bench run therefore repeats a parity check in front of its timed region: the sharded engine against
# (a) the committed golden vectors of the real reference (tests/golden/golden_v1.npz; bit for bit,
# both semantics) and (b) this rank's own single-GPU engine on 26-qubit circuits (QFT, brickwork),
# bit for bit in the default mode and within 1e-12 under math=fast.  The verdict is part of the JSON
# line (sanity.sharded_parity).  Nothing here touches oracle/.
PARITY_QUBITS = 26


def _parity_scripts(n):
    from qcs_b200.workloads import random_circuit_script
    return {"qft": [("x", 1), ("ry", n - 1, 0.3), ("h", n - 2), ("qft",)],
            "brickwork_d10": random_circuit_script(n, 10, seed=5)}


def sharded_parity_reference(Circuit, rank: int, world: int) -> dict:
    """BEFORE the communicator exists: this rank's slice of the single-GPU engine's result."""
    from qcs_b200.workloads import replay
    n = PARITY_QUBITS
    count = (1 << n) // world
    out = {}
    for name, script in _parity_scripts(n).items():
        c = Circuit(n, semantics="corrected")
        replay(c, script)
        full = c.state()
        out[name] = (full[rank * count:(rank + 1) * count].copy(), float(abs(full).max()))
        c.close()
    return out


def sharded_parity_check(Circuit, want: dict, rank: int, world: int, dist, torch) -> dict:
    import numpy as np
    from qcs_b200.workloads import replay
    from tests.golden.cases import CASES
    n = PARITY_QUBITS
    rank_bits = int(math.log2(world))
    res = {"qubits": n, "exact_mismatches": 0, "fast_max_rel_err": 0.0, "golden_cases": 0, "golden_mismatches": 0,
           "remaps": 0, "fused_remaps": 0}
    for name, script in _parity_scripts(n).items():
        shard, scale = want[name]
        for mode in ("exact", "fast"):
            c = Circuit(n, semantics="corrected", math=mode)
            replay(c, script); c.flush()
            st = c.stats()
            got = c.state()
            if mode == "exact":
                res["exact_mismatches"] += int(np.sum(got != shard))
                res["remaps"] += st["remaps"]; res["fused_remaps"] += st["fused_remaps"]
            else:
                res["fast_max_rel_err"] = max(res["fast_max_rel_err"], float(np.abs(got - shard).max() / scale))
            c.close()
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    for name in sorted(CASES):
        nq, script = CASES[name]
        if nq - rank_bits < 2:
            continue
        count = (1 << nq) // world
        for sem in ("reference", "corrected"):
            c = Circuit(nq, semantics=sem)
            vals = replay(c, script)
            key = f"{name}/{sem}"
            bad = int(np.sum(c.state() != golden[key + "/state"][rank * count:(rank + 1) * count]))
            for k, (kind, v) in enumerate(vals):
                bad += int(not np.array_equal(np.asarray(v), golden[f"{key}/out{k}_{kind}"]))
            res["golden_cases"] += 1
            res["golden_mismatches"] += bad
            c.close()
    t = torch.tensor([res["exact_mismatches"], res["golden_mismatches"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    e = torch.tensor([res["fast_max_rel_err"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    res["exact_mismatches"], res["golden_mismatches"] = int(t[0]), int(t[1])
    res["fast_max_rel_err"] = float(e[0])
    res["ok"] = res["exact_mismatches"] == 0 and res["golden_mismatches"] == 0 and res["fast_max_rel_err"] <= 1e-12
    res["what"] = (f"sharded engine on {world} ranks vs each rank's own single-GPU run of a {n}-qubit QFT and a depth-10 "
                   "brickwork circuit (== in the default mode, 1e-12 under math=fast) and vs the committed golden "
                   "vectors of the real reference (==, both semantics)")
    return res


def ncu_traffic(n_local: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the heaviest fused pass, read from the
    committed `ncu --set full` summary of this workload (profiles/, scripts/ncu_summary.py)."""
    import csv
    for name in ("r2_qft30_exact_summary.csv", "r1u_qft30_exact_summary.csv"):
        path = os.path.join(ROOT, "profiles", name)
        if n_local != 30 or not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        try:
            ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
            unit = {"Gbyte": 1e9, "Mbyte": 1e6, "byte": 1.0}.get(rows[1][ir], 1e9)
            launches = [r for r in rows[2:] if "fused_pass" in r[0]]
            heavy = max(launches, key=lambda r: float(r[it]))
            return {"bytes_per_launch": (float(heavy[ir]) + float(heavy[iw])) * unit, "source": f"profiles/{name}",
                    "launch": "heaviest fused pass of the capture"}
        except (ValueError, IndexError):
            continue
    return None


# ----------------------------------------------------------------------------- our arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qcs_b200", choices=["qcs_b200", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the circuit width (default 30 + log2 N)")
    ap.add_argument("--tile-kernel", default=None, help="ldg | tma | tma16 (default: library default)")
    ap.add_argument("--pass-flops", type=float, default=None)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-aux", action="store_true")
    ap.add_argument("--skip-parity", action="store_true", help="N > 1: skip the sharded parity check in front of the timed region")
    ap.add_argument("--aux", action="store_true", help="also run the random-circuit config at N=1 (31 qubits)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: native libraries (NCCL's version banner) that write to
    # file descriptor 1 are sent to stderr, the JSON goes out through a private copy of the descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    warmup = max(3, args.warmup)
    import torch  # first: so that NCCL (dlopen'ed by libqcs_cuda) resolves to torch's copy
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    from qcs_b200 import Circuit, _ffi
    H, C = _ffi.load()

    parity_want = sharded_parity_reference(Circuit, rank, world) if world > 1 and not args.skip_parity else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            assert C.qcs_cuda_dist_unique_id(buf) == 0, _ffi.last_error()
            uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().numpy().tobytes())
        assert C.qcs_cuda_dist_init(rank, world, raw, local_rank) == 0, _ffi.last_error()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if parity_want is not None:
        parity = sharded_parity_check(Circuit, parity_want, rank, world, dist, torch)
        parity_want = None

    n = args.qubits or (30 + int(math.log2(world)))
    gates = qft_gate_count(n)
    kw = {"semantics": "corrected"}
    if args.tile_kernel:
        kw["tile_kernel"] = args.tile_kernel
    if args.pass_flops:
        kw["pass_flops"] = args.pass_flops

    # ---------------- device-resident throughput (`value`) ----------------------------------
    c = Circuit(n, **kw)
    c.set_timing(True)
    for _ in range(warmup):
        c.qft(); c.flush()
    barrier()
    c.reset_stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    c.marker(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c.qft()
        c.flush()
    c.marker(1)
    barrier()
    wall_s = time.perf_counter() - t0
    dev_s = c.marker_elapsed_ms(0, 1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    st = c.stats()
    plan_text = c.describe_plan()
    passes_last_step = c.pass_info()   # the last timed step, pass by pass
    if world > 1:
        t = torch.tensor([dev_s, wall_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall_s = float(t[0]), float(t[1])
    value = gates * world * args.steps / dev_s
    pass_gbs = st["pass_bytes"] / (st["pass_ms"] * 1e-3) / 1e9 if st["pass_ms"] > 0 else 0.0
    # correctness guard inside the bench: the QFT of |0..0> is exactly uniform (corrected semantics);
    # an even number of QFTs of it is checked through the norm instead -- cheap sanity only.
    p0 = c.get_probability(0)
    c.close()

    # ---------------- end to end through the public C API (`e2e`) -----------------------------
    e2e_steps = max(1, args.steps)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    step_ms = []
    phase_s = {"qc_create": 0.0, "gates + qc_find_most_likely_state": 0.0, "qc_get_probability": 0.0, "qc_destroy": 0.0}
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        ce = Circuit(n, **kw)                      # qc_create: buffers from the pool; |0...0> is not written out
        tb = time.perf_counter()
        ce.qft()                                   # 465 qc_h / qc_cphase calls (queued)
        best = ce.find_most_likely_state()         # flush: the fused passes, argmax candidates from the last one, D2H
        tc = time.perf_counter()
        prob = ce.get_probability(12345)           # D2H of one amplitude
        td = time.perf_counter()
        d2h += 16 + 16
        ce.close()                                 # qc_destroy
        te = time.perf_counter()
        for key, dt in zip(phase_s, (tb - ta, tc - tb, td - tc, te - td)):
            phase_s[key] += dt
        step_ms.append(1e3 * (te - ta))
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = gates * world * e2e_steps / e2e_s
    passes_per_step = st["passes"] / args.steps
    # host->device traffic of a step = the gate descriptors, shipped as kernel parameter blocks
    h2d_per_step = int(passes_per_step * C.qcs_cuda_pass_descriptor_bytes())

    aux = None
    strong = None
    if not args.skip_aux:
        # BASELINE config 5 (and its 1-GPU baseline at N = 1: the number the 8-GPU efficiency is divided by)
        aux = random_circuit_aux(Circuit, kw, world, barrier, dist, torch)
        if world > 1:
            strong = strong_scaling_aux(Circuit, kw, world, barrier, dist, torch, args.steps, warmup)
    light = None
    fast = None
    if world == 1 and not args.skip_aux:
        light = light_pass_aux(Circuit, kw, n)
    if not args.skip_aux:
        try:  # an aux report must never cost the headline line
            fast = fast_math_aux(Circuit, kw, n, args.steps, warmup, world, barrier, dist, torch)
            if aux is not None:
                kf = dict(kw); kf["math"] = "fast"
                if strong is not None:
                    strong["math_fast"] = strong_scaling_aux(Circuit, kf, world, barrier, dist, torch, args.steps, warmup)
                fast["random_circuit"] = random_circuit_aux(Circuit, kf, world, barrier, dist, torch)
        except Exception as exc:
            fast = {"failed": repr(exc)}

    if rank != 0:
        if world > 1:
            C.qcs_cuda_dist_finalize()
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"
    n_local = n - int(math.log2(world))
    # FP64 ceiling of the same kernel: the gate arithmetic is separately rounded mul / add (no FMA:
    # bit-exactness), so the ceiling is the issue rate of such instructions, measured by a probe kernel
    fp64_peak = ctypes.c_double(0.0)
    fp64 = None
    if C.qcs_cuda_probe_fp64(ctypes.byref(fp64_peak)) == 0 and st["pass_ms"] > 0:
        ops = st["pass_flops_per_amp"] * 2.0 ** n_local
        fp64 = {"achieved": ops / (st["pass_ms"] * 1e-3) / 1e12, "peak": fp64_peak.value / 1e12,
                "unit": "Tops/s (separately rounded FP64 mul/add, FMA not allowed)",
                "frac": ops / (st["pass_ms"] * 1e-3) / fp64_peak.value,
                "ops_per_amplitude_per_step": st["pass_flops_per_amp"] / args.steps,
                "peak_source": "measured in this run (qcs_cuda_probe_fp64: mul.rn/add.rn chains, no memory traffic)"}
    # Pass by pass (last timed step): a pass costs max(memory time, FP64 issue time), so each pass is held
    # against BOTH ceilings and scored by the larger fraction -- cutting a circuit into more, lighter
    # passes raises the HBM fraction of every pass without making anything faster, this figure does not move.
    per_pass = []
    for k, pi in enumerate(passes_last_step):
        if pi["ms"] <= 0:
            continue
        hbm_frac = 32.0 * 2.0 ** n_local / (pi["ms"] * 1e-3) / 1e9 / peak
        fp64_frac = (pi["flops_per_amp"] * 2.0 ** n_local / (pi["ms"] * 1e-3) / fp64_peak.value) if fp64_peak.value > 0 else None
        per_pass.append({"pass": k, "ms": pi["ms"], "tile_bits": pi["tile_bits"], "segments": pi["segments"],
                         "fp64_ops_per_amplitude": pi["flops_per_amp"], "hbm_frac": hbm_frac, "fp64_frac": fp64_frac,
                         "bound": "fp64" if (fp64_frac or 0) > hbm_frac else "hbm",
                         "frac_of_binding_ceiling": max(hbm_frac, fp64_frac or 0.0)})
    tot_ms = sum(q["ms"] for q in per_pass)
    binding = (sum(q["frac_of_binding_ceiling"] * q["ms"] for q in per_pass) / tot_ms) if tot_ms > 0 else None
    traffic = ncu_traffic(n_local)
    line = {
        "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * dev_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"{n}-qubit QFT (qc_quantum_fourier_transform), {gates} gates, "
                        f"{16 * 2 ** n_local / 2 ** 30:.0f} GiB of amplitudes per GPU, corrected semantics",
            "qubits": n, "gates_per_step": gates, "parallelism": f"shard{world}",
            "l2": "inputs larger than L2 (16 GiB shard vs 126 MB L2), no flush needed",
            "value_definition": "gates x shards / device time (CUDA events on the engine stream, max over ranks)",
            "tile_kernel": args.tile_kernel or "default", "semantics": "corrected",
        },
        "wall_ms_per_step": 1e3 * wall_s / args.steps,
        "passes_per_step": passes_per_step,
        "gates_per_pass": gates / passes_per_step if passes_per_step else None,
        "remaps_per_step": st["remaps"] / args.steps,
        "fused_remaps_per_step": st["fused_remaps"] / args.steps,
        "out_of_place_remaps_per_step": st["out_of_place_remaps"] / args.steps,
        "fused_remap_pass_ms": (st["fused_remap_pass_ms"] / st["fused_remaps"]) if st["fused_remaps"] else None,
        "roofline": {
            "bound": "hbm", "kernel": "fused_pass (tile kernel)",
            "achieved": pass_gbs, "peak": peak, "unit": "GB/s", "frac": pass_gbs / peak,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": 32.0 * 2 ** n_local,
            "avg_launch_ms": st["pass_ms"] / st["passes"] if st["passes"] else None,
            "traffic": traffic["bytes_per_launch"] if traffic else None,
            "traffic_source": traffic,
            "share_of_step": (st["pass_ms"] * 1e-3) / dev_s if dev_s > 0 else None,
            "co_limiter_fp64": fp64,
            "per_pass": per_pass,
            "time_weighted_frac_of_binding_ceiling": binding,
        },
        "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": h2d_per_step,
                "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps,
                "host_ms_per_step_by_call": {k2: 1e3 * v / e2e_steps for k2, v in phase_s.items()},
                "host_ms_of_each_step": [round(v, 2) for v in step_ms],
                "what": "qc_create + qc_quantum_fourier_transform + qc_find_most_likely_state + "
                        "qc_get_probability + qc_destroy through libqcs.so"},
        "gpu_launches": st["kernel_launches"],
        "clocks": clocks,
        "sanity": {"p(|0>) after the timed QFTs": p0, "sharded_parity": parity},
    }
    if strong:
        line["aux_strong_scaling_qft30"] = strong
    if aux:
        line["aux_random_circuit"] = aux
    if light:
        line["aux_bandwidth_bound_passes"] = light
    if fast:
        line["aux_fast_math"] = fast
    if world == 1 and not args.skip_cpu_baseline:
        try:
            line["cpu_baseline"] = {k: v for k, v in run_reference_sample(n, 20.0).items()
                                    if k in ("value", "unit", "cores", "kind", "sample", "modes", "host_cores", "cpu_model")}
        except Exception as exc:  # the baseline is a report, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": 0, "kind": "reference",
                                    "sample": f"failed: {exc}"}
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        C.qcs_cuda_dist_finalize()
        dist.destroy_process_group()


def fast_math_aux(Circuit, kw, n, steps, warmup, world, barrier, dist, torch) -> dict:
    """The headline workload under the opt-in math=fast mode (fused multiply-adds, controlled-phase
    fans collapsed into one factor per thread): amplitudes within 1e-12 of the reference instead of
    bit-exact (tests/test_gpu_fast_math.py), so it is reported beside the headline, not as it."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else HBM_FALLBACK_GBS
    k = dict(kw); k["math"] = "fast"
    c = Circuit(n, **k)
    c.set_timing(True)
    for _ in range(warmup):
        c.qft(); c.flush()
    barrier()
    c.reset_stats()
    c.marker(0)
    for _ in range(steps):
        c.qft(); c.flush()
    c.marker(1)
    barrier()
    dev_s = c.marker_elapsed_ms(0, 1) * 1e-3
    st = c.stats()
    c.close()
    if world > 1:
        t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t[0])
    gates = qft_gate_count(n)
    gbs = st["pass_bytes"] / (st["pass_ms"] * 1e-3) / 1e9
    return {"mode": "math=fast, parity bar 1e-12 relative (not bit-exact)",
            "value": gates * world * steps / dev_s, "unit": "gates/s", "ms_per_step": 1e3 * dev_s / steps,
            "remaps_per_step": st["remaps"] / steps,
            "passes_per_step": st["passes"] / steps,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "avg_launch_ms": st["pass_ms"] / st["passes"]},
            "fp64_instructions_per_amplitude_per_step": st["pass_flops_per_amp"] / steps}


def strong_scaling_aux(Circuit, kw, world, barrier, dist, torch, steps, warmup) -> dict:
    """the 30-qubit QFT (fixed total work) sharded over the N GPUs of this run."""
    n = 30
    c = Circuit(n, **kw)
    c.set_timing(True)
    for _ in range(warmup):
        c.qft(); c.flush()
    barrier()
    c.reset_stats()
    c.marker(0)
    for _ in range(steps):
        c.qft(); c.flush()
    c.marker(1)
    barrier()
    dev_s = c.marker_elapsed_ms(0, 1) * 1e-3
    st = c.stats()
    c.close()
    t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t[0])
    return {"workload": f"30-qubit QFT on {world} GPUs ({16 // world} GiB per GPU), strong scaling",
            "gates_per_s": qft_gate_count(n) * steps / dev_s, "ms_per_step": 1e3 * dev_s / steps,
            "passes_per_step": st["passes"] / steps, "remaps_per_step": st["remaps"] / steps,
            "fused_remap_pass_ms": (st["fused_remap_pass_ms"] / st["fused_remaps"]) if st["fused_remaps"] else None}


def light_pass_aux(Circuit, kw, n) -> dict:
    """The same kernel on passes that carry little arithmetic: how close a fused pass gets to the
    HBM roofline when FP64 issue is not the limiter (the headline QFT passes fuse ~190 separately
    rounded FP64 operations per amplitude; these fuse 4 - 48)."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else HBM_FALLBACK_GBS
    out = {}
    cases = {"one_H_per_pass": ([("h", 8)], {}),
             "H_on_every_qubit": ([("h", q) for q in range(n)], {}),
             "QFT_cut_at_48_flops_per_pass": ([("qft",)], {"pass_flops": 48.0})}
    from qcs_b200.workloads import replay
    for name, (script, extra) in cases.items():
        k = dict(kw); k.update(extra)
        c = Circuit(n, **k)
        c.set_timing(True)
        replay(c, script); c.flush(); c.reset_stats()
        for _ in range(3):
            replay(c, script); c.flush()
        st = c.stats()
        c.close()
        gbs = st["pass_bytes"] / (st["pass_ms"] * 1e-3) / 1e9
        out[name] = {"passes": st["passes"] // 3, "gates": st["gates_submitted"] // 3,
                     "ms": st["pass_ms"] / 3, "GBps_per_pass": gbs, "frac_of_hbm_peak": gbs / peak}
    return out


def random_circuit_aux(Circuit, kw, world, barrier, dist, torch) -> dict:
    """BASELINE config 5: H/CNOT/RZ brickwork, depth 40, 31+log2(N) qubits (32 GiB shards)."""
    from qcs_b200.workloads import random_circuit_script, replay
    n = 31 + int(math.log2(world))
    script = random_circuit_script(n, 40)
    c = Circuit(n, **kw)
    c.set_timing(True)
    replay(c, script[: len(script) // 8]); c.flush()   # warm-up slice
    barrier()
    c.reset_stats()
    c.marker(0)
    replay(c, script)
    c.flush()
    c.marker(1)
    barrier()
    dev_s = c.marker_elapsed_ms(0, 1) * 1e-3
    st = c.stats()
    c.close()
    if world > 1:
        t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t[0])
    return {"workload": f"{n}-qubit random circuit (H/RZ/CNOT brickwork, depth 40, {len(script)} gates)",
            "gates_per_s": len(script) / dev_s, "seconds": dev_s, "passes": st["passes"],
            "amplitude_gate_updates_per_s_per_gpu": len(script) * 2.0 ** n / dev_s / world,
            "remaps": st["remaps"], "fused_remaps": st["fused_remaps"], "out_of_place_remaps": st["out_of_place_remaps"],
            "fused_remap_pass_ms_avg": (st["fused_remap_pass_ms"] / st["fused_remaps"]) if st["fused_remaps"] else None,
            "position_pairs_traded": st["remaps"], "carrying_passes_with_2_or_3_pairs": st["multi_remaps"],
            "nvlink_GBps_per_direction_in_fused_passes": (st["fused_remap_bytes"] / (st["fused_remap_pass_ms"] * 1e-3) / 1e9
                                                          if st["fused_remap_pass_ms"] else None),
            "nvlink_frac_of_measured_peer_copy_770GBps": (st["fused_remap_bytes"] / (st["fused_remap_pass_ms"] * 1e-3) / 1e9 / 770.0
                                                          if st["fused_remap_pass_ms"] else None),
            "plain_pass_ms_avg": ((st["pass_ms"] - st["fused_remap_pass_ms"]) / max(1, st["passes"] - st["fused_remaps"])),
            "exchange_ms": st["exchange_ms"], "pass_ms": st["pass_ms"],
            "pass_GBps": st["pass_bytes"] / (st["pass_ms"] * 1e-3) / 1e9 if st["pass_ms"] else None,
            "exchange_GBps_per_direction": (st["exchange_bytes"] / (st["exchange_ms"] * 1e-3) / 1e9
                                            if st["exchange_ms"] else None)}


if __name__ == "__main__":
    main()
