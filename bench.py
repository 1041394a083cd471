#!/usr/bin/env python
"""bench.py — two-site gates/s (and BP-sweep ms) of the BP simple-update path on L×L TFIM at bond
dimension χ (BASELINE.json metric), B200 vs the CPU reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--L 16] [--chi 32]

A *step* is one Trotter layer of examples/2dIsing_dynamics.jl (Rx, Rz on every vertex, then the Rzz
colour groups) applied with `apply_gates` — i.e. 2·|V| one-site gates, |E| two-site gates and
(colours + 1) BP refreshes.  The state is evolved from the all-↑ product state for `--prep` layers
first so that the timed layers run at saturated bond dimension χ (reported in `config`).

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# load every kernel of the library at context creation: with lazy loading the first launch of each
# kernel variant inside the timed region would pay its module-load stall
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "two_site_gates_per_sec"
UNIT = "gates/s"


def tfim_layer(tq, g, dt=0.25, hx=1.0, hz=0.8, J=0.5):
    """examples/2dIsing_dynamics.jl:12-28."""
    layer = [("Rx", [v], 2 * hx * dt) for v in g.vertices()]
    layer += [("Rz", [v], 2 * hz * dt) for v in g.vertices()]
    groups = tq.edge_color(g, 4)
    for grp in groups:
        layer += [("Rzz", list(pair), 2 * J * dt) for pair in grp]
    return layer, len(groups)


def build_workload(tq, args):
    """Graph, one circuit layer, number of colour groups, apply kwargs and a description for the BASELINE.json configs:
    tfim2d   — L×L open square lattice, layer of examples/2dIsing_dynamics.jl:12-28 (configs 1, 2, 5 and the χ=64 target)
    heavyhex — IBM-Eagle 127-qubit heavy-hex, kicked-Ising layer of examples/heavyhexIsing_dynamics.jl:12-26 (config 3)
    cubic3d  — L×L×L periodic cubic lattice, layer of examples/3dIsing_dynamics.jl:15-26 (config 4)"""
    chi = args.chi
    if args.workload == "tfim2d":
        g = tq.named_grid((args.L, args.L))
        layer, ncol = tfim_layer(tq, g)
        kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
        name = (f"{args.L}x{args.L} square-lattice TFIM (examples/2dIsing_dynamics.jl constants), maxdim={chi}, cutoff=1e-10, ComplexF32")
        centre = (args.L // 2 + 1, args.L // 2 + 1)
        degs = (4, 4)
    elif args.workload == "heavyhex":
        g = tq.eagle_heavy_hex()
        groups = tq.edge_color(g, 3)
        layer = [("Rx", [v], 0.4) for v in g.vertices()]
        for grp in groups:
            layer += [("Rzz", list(pair), np.pi / 2) for pair in grp]
        ncol = len(groups)
        kw = dict(maxdim=chi, cutoff=1e-12, normalize_tensors=True)
        name = f"IBM-Eagle heavy-hex 127 qubits, kicked Ising (examples/heavyhexIsing_dynamics.jl: Rx(0.4), Rzz(pi/2)), maxdim={chi}, cutoff=1e-12, ComplexF32"
        centre = g.center()[0]
        degs = (3, 2)
    elif args.workload == "cubic3d":
        g = tq.named_grid((args.L, args.L, args.L), periodic=True)
        groups = tq.edge_color(g, 6)
        h, J, dt = -1.0, -1.0, 0.04
        layer = [("Rz", [v], h * dt) for v in g.vertices()]
        for grp in groups:
            layer += [("Rxx", list(pair), 2 * J * dt) for pair in grp]
        layer += [("Rz", [v], h * dt) for v in g.vertices()]
        ncol = len(groups)
        kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
        name = f"{args.L}x{args.L}x{args.L} periodic cubic Ising (examples/3dIsing_dynamics.jl constants), maxdim={chi}, cutoff=1e-10, ComplexF32"
        centre = g.center()[0]
        degs = (6, 6)
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    n_two = sum(1 for gt in layer if len(gt[1]) == 2)
    n_one = len(layer) - n_two
    name += f", one Trotter layer per step = {n_two} two-site + {n_one} one-site gates + {ncol+1} BP refreshes"
    return g, layer, ncol, kw, name, centre, degs


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, sampled during the timed region."""

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm: the NumPy/OpenBLAS oracle (the Julia reference cannot run in this image)
# ---------------------------------------------------------------------------------------------

def _blas_threads_all():
    """Use every host core for BLAS (torchrun exports OMP_NUM_THREADS=1); returns (context manager, thread count)."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n), n
    except Exception:
        import contextlib
        return contextlib.nullcontext(), n


def _hub_patch(tq, orc, z1, z2, chi, seed):
    """Two adjacent vertices of degrees (z1, z2) whose other neighbours are leaves: the gate on the hub edge and the
    message leaving hub 1 cost what an interior gate / message of the workload's lattice costs."""
    vs = ["a", "b"] + [f"a{i}" for i in range(z1 - 1)] + [f"b{i}" for i in range(z2 - 1)]
    es = [("a", "b")] + [("a", f"a{i}") for i in range(z1 - 1)] + [("b", f"b{i}") for i in range(z2 - 1)]
    gp = tq.NamedGraph(vs, es)
    rng = np.random.default_rng(seed)
    c = orc.random_state(gp.nv, gp.edge_uv(), 2, chi, np.complex64, seed=seed)
    for (u, v) in c.directed_edges():
        w = (rng.standard_normal((chi, chi)) + 1j * rng.standard_normal((chi, chi))).astype(np.complex64)
        m = w @ w.conj().T + np.eye(chi, dtype=np.complex64)
        c.msg[(u, v)] = (m / m.sum()).astype(np.complex64)
    return gp, c


def reference_sweeps_per_refresh(tq, orc):
    """BP sweeps the reference's own `update` needs per refresh in steady state, measured with the oracle on a system it can
    evolve (4×4 TFIM, χ = 8, ComplexF32 defaults maxiter 25 / tolerance 1e-5, layers 5 and 6 of the evolution)."""
    g = tq.named_grid((4, 4))

    class A:
        pass
    layer, ncol = tfim_layer(tq, g)
    nverts, verts, mats = tq.circuit_arrays(layer, g)
    mc = mats.view(np.complex128)
    gm, off = [], 0
    for n in nverts:
        k = 4 ** int(n)
        gm.append(mc[off:off + k].reshape(2 ** int(n), 2 ** int(n)))
        off += k
    gv = [[int(x) for x in v[:n]] for v, n in zip(verts, nverts)]
    seq = [(g.index[a], g.index[b]) for a, b in tq.bipartite_edge_sequence(g)]
    c = orc.product_state(g.nv, g.edge_uv(), [(1.0, 0.0)] * g.nv, np.complex64)
    per = []
    for l in range(6):
        c, _, reps = orc.apply_gates(c, gm, gv, seq, dict(maxdim=8, cutoff=1e-10, normalize_tensors=True),
                                     dict(maxiter=25, tolerance=1e-5))
        per.append(sum(r["niter"] for r in reps) / len(reps))
    return float(np.mean(per[-2:]))


def cpu_reference_sample(args, sweeps_per_refresh, budget_s=20.0, seed=1234):
    """Time the oracle (the NumPy/LAPACK restatement of the reference's CPU path; Julia is not installed here) on a
    bounded sample of the same workload — interior two-site gates and interior message updates at bond dimension χ on
    a two-hub patch with the workload's vertex degrees — and extrapolate to one layer = n_two gates + (colours + 1)
    refreshes × sweeps_per_refresh × 2|E| messages.  Returns (gates/s, description, ms per BP sweep, threads)."""
    import tnqs_b200 as tq
    from oracle import tnqs_oracle as orc
    ctx, nthreads = _blas_threads_all()
    g, layer, ncol, kw, name, centre, degs = build_workload(tq, args)
    n_two = sum(1 for gt in layer if len(gt[1]) == 2)
    chi = args.chi
    with ctx:
        gp, c = _hub_patch(tq, orc, degs[0], degs[1], chi, seed)
        a, b = gp.index["a"], gp.index["b"]
        gname = [gt for gt in layer if len(gt[1]) == 2][0][0]
        gate = tq.gate_matrix(gname, 2, 0.25)
        t_gate, n_gate = 0.0, 0
        t0 = time.perf_counter()
        while n_gate < 1 or (time.perf_counter() - t0 < budget_s / 2 and n_gate < 8):
            cc = c.copy()
            t1 = time.perf_counter()
            orc.apply_gate(cc, gate, [a, b], maxdim=chi, cutoff=kw["cutoff"], normalize_tensors=True)
            t_gate += time.perf_counter() - t1
            n_gate += 1
        t_msg, n_msg = 0.0, 0
        t0 = time.perf_counter()
        while n_msg < 1 or (time.perf_counter() - t0 < budget_s / 2 and n_msg < 64):
            t1 = time.perf_counter()
            orc.updated_message(c, a, b)
            t_msg += time.perf_counter() - t1
            n_msg += 1
    tg, tm = t_gate / n_gate, t_msg / n_msg
    sweeps = sweeps_per_refresh * (ncol + 1)
    layer_s = n_two * tg + sweeps * 2 * g.ne * tm
    desc = (f"oracle (NumPy/OpenBLAS, {nthreads} BLAS threads), complex64: {n_gate} interior two-site gates between vertices of "
            f"degree {degs} ({tg*1e3:.1f} ms each) + {n_msg} interior message updates ({tm*1e3:.2f} ms each) at chi={chi}, "
            f"extrapolated to one layer = {n_two} gates + {sweeps:.1f} BP sweeps x {2*g.ne} messages")
    return n_two / layer_s, desc, 2 * g.ne * tm * 1e3, nthreads, name


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tnqs_b200 as tq
    from oracle import tnqs_oracle as orc
    ctx, _ = _blas_threads_all()
    with ctx:
        spr = args.ref_bp_sweeps if args.ref_bp_sweeps > 0 else reference_sweeps_per_refresh(tq, orc)
    vals, walls = [], []
    desc, nthreads, sweep_ms, name = "", 1, 0.0, ""
    for i in range(args.warmup + args.steps):
        tw = time.perf_counter()
        v, desc, sweep_ms, nthreads, name = cpu_reference_sample(args, spr, budget_s=args.ref_budget)
        if i >= args.warmup:
            vals.append(v)
            walls.append(time.perf_counter() - tw)
    val = float(np.mean(vals)) if vals else 0.0
    g, layer, ncol, kw, name, centre, degs = build_workload(tq, args)
    n_two = sum(1 for gt in layer if len(gt[1]) == 2)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(walls)) if walls else None,  # wall time of one bounded sample (what a step runs)
        "ms_per_layer_extrapolated": (1e3 * n_two / val) if val else None,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": {"workload": name, "bp_schedule": args.schedule,
                   "bp_sweeps_per_layer": spr * (ncol + 1),
                   "bp_sweeps_per_refresh_measured": spr,
                   "how": "each step times a bounded sample of the layer on the host cores and extrapolates (a full chi>=32 layer takes "
                          "the CPU path tens of minutes); sweeps per refresh measured with the oracle's own update on a 4x4 chi=8 TFIM evolution"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": desc},
        "bp_sweep_ms": sweep_ms,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import tnqs_b200 as tq

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tnqs_b200 has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = np.complex64
    chi = args.chi
    g, layer, ncol, kw, wname, centre, degs = build_workload(tq, args)
    n_two = sum(1 for gt in layer if len(gt[1]) == 2)
    seq = tq.bipartite_edge_sequence(g) if args.schedule == "bipartite" else tq.forest_cover_edge_sequence(g)
    bp = dict(maxiter=25, tolerance=1e-5, edge_sequence=seq)  # default_bp_update_kwargs for ComplexF32
    if args.random_state:
        # BASELINE config 5 / large chi: synthetic random TNS with every bond = χ (iid normal entries keyed by seed 1234, each
        # tensor scaled to unit Frobenius norm), generated on the device (16×16 at χ=64 is 53 GB), then one BP update
        psi = tq.random_bpc_on_device(dtype, g, bond_dimension=chi, seed=1234, device=local,
                                      shard_fn=(tq.shard if world > 1 else None))
        psi = tq.update(psi, inplace=True, **bp)
    else:
        psi = tq.BeliefPropagationCache(tq.tensornetworkstate(dtype, lambda v: "↑", g, "S=1/2"), device=local)
        if world > 1:
            tq.shard(psi)  # contiguous vertex blocks, one per rank; messages / Gram matrices travel over NCCL
    obs = ("Z", [centre])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_prep = time.perf_counter()
    nprep = args.prep if args.prep is not None else (0 if args.random_state else 15)
    for _ in range(nprep):
        psi, _ = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True)
    for _ in range(args.warmup):  # same call as the timed steps (functional copy): warms the memory pools of that path
        psi, _ = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=args.inplace)
        tq.expect(psi, obs)
    t_prep = time.perf_counter() - t_prep
    bd = psi.bond_dims()

    nverts, verts, mats = tq.circuit_arrays(layer, g)
    h2d = int(nverts.nbytes + verts.nbytes + mats.nbytes)
    psi.stats(reset=True)
    sampler = ClockSampler(local)
    if not args.no_sampler:
        sampler.start()
    sweeps, e2e_s, dev_ms, zs = 0, 0.0, 0.0, []
    maxerr = 0.0
    st = {"bp_ms": 0.0, "su_ms": 0.0, "bp_sweeps": 0, "kernel_launches": 0}
    step_ms = []
    if args.cuda_profiler:
        torch.cuda.profiler.start()
    sync_all()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=args.inplace)  # public API, host in/out
        z = tq.expect(psi, obs)
        e2e_s += time.perf_counter() - t0
        zs.append(float(np.real(z)))
        maxerr = max(maxerr, float(errs.max()))
        sweeps += sum(r["niter"] for r in psi.last_bp_reports)
        s1 = psi.stats(reset=args.inplace)  # the returned cache is a fresh clone: its counters cover exactly this call
        for k in st:
            st[k] += s1[k]
        step_ms.append(s1["bp_ms"] + s1["su_ms"])
    sync_all()
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    dev_ms = st["bp_ms"] + st["su_ms"]  # CUDA events on the engine stream around every apply_gates call
    if world > 1:  # device time and wall time: max over ranks
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(tt[0]), float(tt[1])
    d2h = int(8 * len(nverts) + 16)
    value = n_two * args.steps / (dev_ms * 1e-3)
    e2e = n_two * args.steps / e2e_s
    sweeps_per_layer = sweeps / args.steps
    bp_sweep_ms = st["bp_ms"] / max(1, st["bp_sweeps"])

    # roofline of the dominant kernel family, timed live with CUDA events (profiling mode brackets
    # every launch group with events on the launching stream; done on one extra, untimed layer)
    psi.set_profiling(True)
    psi.stats(reset=True)
    psi2, _ = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True)
    sp = psi2.stats()
    psi.set_profiling(False)
    pk = peaks()
    fams = {}
    for nm, ms, by, fl, nl in (("mode_product", sp["mode_ms"], sp["mode_bytes"], sp["mode_flops"], sp["mode_launches"]),
                               ("gram", sp["gram_ms"], sp["gram_bytes"], sp["gram_flops"], sp["gram_launches"])):
        gbs = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        fams[nm] = {"ms_one_layer": ms, "algorithmic_GB": by / 1e9, "GBps": gbs, "frac_of_hbm_peak": gbs / pk["hbm"],
                    "algorithmic_tflops": fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0, "launches": int(nl)}
    fams["jacobi_cholesky_small"] = {"ms_one_layer": sp["small_ms"]}
    # dominant tensor-streaming kernel family; its algorithmic bytes (every tensor read once and
    # written once) over its CUDA-event time, against the measured HBM copy peak
    fam = max(("mode_product", "gram"), key=lambda k: fams[k]["ms_one_layer"])
    kernel_names = {"mode_product": "tc2_mode_kernel (TMA-fed tcgen05, 3xTF32) + tc_mode_kernel fallback",
                    "gram": "tc_gram_kernel (tcgen05 3xTF32, BP closing contraction) + gram_dmma_kernel (fp64 tensor core, simple update)"}
    traffic, traffic_note = None, "no ncu capture of this workload committed"
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(f"{args.workload}:{args.L}:{chi}", {}).get(fam)
        if tj:
            traffic, traffic_note = tj["dram_bytes_per_launch"], tj["note"]
    roof = {"bound": "hbm", "achieved": fams[fam]["GBps"], "peak": pk["hbm"], "unit": "GB/s", "frac": fams[fam]["frac_of_hbm_peak"],
            "traffic": traffic, "traffic_note": traffic_note, "kernel": kernel_names[fam],
            "launches": fams[fam]["launches"], "peak_source": pk["source"],
            "algorithmic_tflops": fams[fam]["algorithmic_tflops"],
            "tc_launches": int(sp["tc_launches"]), "tma_launches": int(sp["tma_launches"]),
            "families": fams,
            "other_peaks_measured": {"tf32_tcgen05_tflops": 1114.3, "fp32_ffma_tflops": 71.5, "fp64_dmma_tflops": 37.0,
                                     "source": "profiles/r2a_tma_probe_and_peaks.txt (tools/tma_probe.cu on this pool's B200)"}}

    # the reference's default schedule (forest_cover_edge_sequence, beliefpropagationcache.jl:27-29): one extra untimed sweep
    extras = {}
    if args.extras and world == 1:
        fseq = tq.forest_cover_edge_sequence(g)
        psi2.stats(reset=True)
        tq.update(psi2, inplace=True, maxiter=1, tolerance=None, edge_sequence=fseq)
        extras["bp_sweep_ms_forest_schedule"] = psi2.stats()["bp_ms"]
        psi2.stats(reset=True)
        tq.update(psi2, inplace=True, maxiter=1, tolerance=None, edge_sequence=tq.bipartite_edge_sequence(g))
        extras["bp_sweep_ms_bipartite_schedule"] = psi2.stats()["bp_ms"]

    cpu = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if not args.no_cpu and world == 1:
        from oracle import tnqs_oracle as orc
        ctx, _ = _blas_threads_all()
        with ctx:
            spr = reference_sweeps_per_refresh(tq, orc)
        v, desc, cpu_sweep_ms, nthreads, _ = cpu_reference_sample(args, spr, budget_s=args.ref_budget)
        cpu = {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": desc, "bp_sweep_ms": cpu_sweep_ms,
               "bp_sweeps_per_refresh_measured": spr}

    state_gb = sum(2 * int(np.prod([2] + [bd[e] for e, _ in g.incident[i]])) * 4 for i in range(g.nv)) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c64", "data": "synthetic",
        "config": {"workload": wname,
                   "prep_layers": nprep, "initial_state": ("random TNS generated on the device, all bonds = chi, seed 1234" if args.random_state else "product state |↑…↑>"),
                   "bond_dim_min_mean_max": [int(bd.min()), float(bd.mean()), int(bd.max())],
                   "bp_schedule": args.schedule, "bp_sweeps_per_layer": sweeps_per_layer,
                   "functional_copy": not args.inplace,
                   "sharding": ("none" if world == 1 else f"contiguous vertex blocks over {world} ranks; NCCL: broadcast of level messages + Gram matrices, all-gather of the per-gate factorisation results (gate k solved on rank k mod {world})"),
                   "l2": "inputs larger than L2 (state %.2f GB)" % state_gb if state_gb > 0.13 else "state %.3f GB fits L2: an L2 flush (256 MB write) is not applied, the step streams ~%d x the state through temporaries" % (state_gb, 20),
                   "max_trunc_err": maxerr, "sz_center": zs, "step_ms": step_ms},
        "bp_sweep_ms": bp_sweep_ms,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "extras": extras,
        "prep_seconds": t_prep,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tfim2d", choices=["tfim2d", "heavyhex", "cubic3d"],
                    help="tfim2d: BASELINE configs 1/2/5 and the chi=64 target; heavyhex: config 3 (Eagle 127); cubic3d: config 4 (periodic LxLxL)")
    ap.add_argument("--L", type=int, default=None, help="linear lattice size (default: 16 for tfim2d, 6 for cubic3d = BASELINE config 4)")
    ap.add_argument("--chi", type=int, default=32)
    ap.add_argument("--prep", type=int, default=None, help="untimed layers before warm-up (default: 15 from the product state, 0 with --random-state)")
    ap.add_argument("--random-state", action="store_true", help="start from a synthetic random TNS with all bonds = chi (BASELINE config 5) instead of evolving the product state")
    ap.add_argument("--schedule", default="bipartite", choices=["bipartite", "forest"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--inplace", action="store_true", help="diagnostic: mutate the cache instead of the reference's functional copy")
    ap.add_argument("--no-sampler", action="store_true", help="diagnostic: do not poll nvidia-smi during the timed region")
    ap.add_argument("--cuda-profiler", action="store_true", help="cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)")
    ap.add_argument("--ref-budget", type=float, default=20.0)
    ap.add_argument("--ref-bp-sweeps", type=float, default=-1.0, help="reference arm: BP sweeps per refresh (default: measured with the oracle)")
    ap.add_argument("--extras", action="store_true", help="also time one BP sweep with the reference's default forest-cover schedule")
    args = ap.parse_args()
    if args.L is None:
        args.L = 6 if args.workload == "cubic3d" else 16
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
