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

def cpu_reference_sample(L, chi, bp_iters_per_layer, ncolors, budget_s=20.0, seed=1234):
    """Time the oracle on a bounded sample of the same workload: interior two-site gates and message
    updates on a random χ-saturated complex64 TNS patch, extrapolated to one layer with the BP sweep
    count the GPU run needed.  Returns (gates/s, description, ms per BP sweep)."""
    import tnqs_b200 as tq
    from oracle import tnqs_oracle as orc
    try:
        from threadpoolctl import threadpool_info
        nthreads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        nthreads = os.cpu_count() or 1
    g = tq.named_grid((L, L))
    # a 4×4 patch has interior vertices of full degree 4; interior-gate / interior-message costs are
    # what dominate the L×L lattice
    gp = tq.named_grid((4, 4))
    rng = np.random.default_rng(seed)
    c = orc.random_state(gp.nv, gp.edge_uv(), 2, chi, np.complex64, seed=seed)
    for (u, v) in c.directed_edges():
        n = chi
        w = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
        m = w @ w.conj().T + np.eye(n, dtype=np.complex64)
        c.msg[(u, v)] = (m / m.sum()).astype(np.complex64)
    a, b = gp.index[(2, 2)], gp.index[(3, 2)]
    gate = tq.gate_matrix("Rzz", 2, 0.25)
    t_gate, n_gate = 0.0, 0
    t0 = time.perf_counter()
    while n_gate < 1 or (time.perf_counter() - t0 < budget_s / 2 and n_gate < 8):
        cc = c.copy()
        t1 = time.perf_counter()
        orc.apply_gate(cc, gate, [a, b], maxdim=chi, cutoff=1e-10, normalize_tensors=True)
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
    layer_s = g.ne * tg + bp_iters_per_layer * 2 * g.ne * tm
    desc = (f"oracle (NumPy/OpenBLAS, {nthreads} BLAS threads), complex64: {n_gate} interior two-site gates "
            f"({tg*1e3:.1f} ms each) + {n_msg} interior message updates ({tm*1e3:.2f} ms each) at chi={chi}, "
            f"extrapolated to one {L}x{L} layer = {g.ne} gates + {bp_iters_per_layer:.1f} BP sweeps x {2*g.ne} messages")
    return g.ne / layer_s, desc, 2 * g.ne * tm * 1e3, nthreads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    desc, nthreads, sweep_ms = "", 1, 0.0
    for i in range(args.warmup + args.steps):
        v, desc, sweep_ms, nthreads = cpu_reference_sample(args.L, args.chi, args.ref_bp_sweeps, 4,
                                                           budget_s=args.ref_budget)
        if i >= args.warmup:
            vals.append(v)
    val = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (args.L * (args.L - 1) * 2) / val,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": {"workload": f"{args.L}x{args.L} square-lattice TFIM layer, chi={args.chi}, ComplexF32",
                   "assumed_bp_sweeps_per_layer": args.ref_bp_sweeps},
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
    L, chi = args.L, args.chi
    g = tq.named_grid((L, L))
    layer, ncol = tfim_layer(tq, g)
    n_two = g.ne
    seq = tq.bipartite_edge_sequence(g) if args.schedule == "bipartite" else tq.forest_cover_edge_sequence(g)
    kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
    bp = dict(maxiter=25, tolerance=1e-5, edge_sequence=seq)  # default_bp_update_kwargs for ComplexF32
    if args.random_state:
        # BASELINE config 5: synthetic random TNS with every bond = χ (iid normal entries, seed 1234, each tensor
        # scaled to unit Frobenius norm), one BP update; no evolution from the product state
        tns = tq.random_tensornetworkstate(dtype, g, bond_dimension=chi, seed=1234)
        for v in list(tns.tensors):
            tns.tensors[v] = (tns.tensors[v] / np.linalg.norm(tns.tensors[v])).astype(dtype)
        psi = tq.BeliefPropagationCache(tns, device=local)
        del tns
    else:
        psi = tq.BeliefPropagationCache(tq.tensornetworkstate(dtype, lambda v: "↑", g, "S=1/2"), device=local)
    if world > 1:
        tq.shard(psi)  # row strips of the lattice, one per rank; messages / Gram matrices travel over NCCL
    if args.random_state:
        psi = tq.update(psi, inplace=True, **bp)
    obs = ("Z", [(L // 2 + 1, L // 2 + 1)])

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
        tq.expect(psi, ("Z", [(L // 2 + 1, L // 2 + 1)]))
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
    t_region = time.perf_counter()
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
    t_region = time.perf_counter() - t_region
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
    # dominant tensor-streaming kernel family; its algorithmic bytes (every tensor read once and
    # written once) over its CUDA-event time, against the measured HBM copy peak
    fam = max((("mode_product", sp["mode_ms"]), ("gram", sp["gram_ms"])), key=lambda x: x[1])
    by = sp["mode_bytes"] if fam[0] == "mode_product" else sp["gram_bytes"]
    fl = sp["mode_flops"] if fam[0] == "mode_product" else sp["gram_flops"]
    nl = sp["mode_launches"] if fam[0] == "mode_product" else sp["gram_launches"]
    ach = by / (fam[1] * 1e-3) / 1e9 if fam[1] > 0 else 0.0
    roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
            "traffic": None, "traffic_note": "ncu --set full of the 8x8 run (profiles/r1s_prof_tc_mode_raw.csv): a tc_mode launch moves 0.758 GB read + 0.856 GB written in DRAM; its output (= input) size is 0.856 GB, i.e. no re-reads (reads 11 % below the algorithmic bytes: L2 hits on lines the previous launch wrote)",
            "kernel": ("tc_mode_kernel (tcgen05 3xTF32)" if fam[0] == "mode_product" else "tc_gram_kernel (tcgen05 3xTF32)"),
            "launches": int(nl), "peak_source": pk["source"],
            "algorithmic_tflops": fl / (fam[1] * 1e-3) / 1e12 if fam[1] > 0 else 0.0,
            "tc_launches": int(sp["tc_launches"]),
            "family_ms_one_layer": {"mode_product": sp["mode_ms"], "gram": sp["gram_ms"], "jacobi": sp["small_ms"]}}

    cpu = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if not args.no_cpu and world == 1:
        v, desc, cpu_sweep_ms, nthreads = cpu_reference_sample(L, chi, sweeps_per_layer, ncol, budget_s=args.ref_budget)
        cpu = {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": desc, "bp_sweep_ms": cpu_sweep_ms}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c64", "data": "synthetic",
        "config": {"workload": f"{L}x{L} square-lattice TFIM (examples/2dIsing_dynamics.jl constants), maxdim={chi}, "
                               f"cutoff=1e-10, ComplexF32, one Trotter layer per step = {n_two} two-site + {2*g.nv} one-site "
                               f"gates + {ncol+1} BP refreshes",
                   "prep_layers": nprep, "initial_state": ("random TNS, all bonds = chi, seed 1234" if args.random_state else "product state |↑…↑>"), "bond_dim_min_mean_max": [int(bd.min()), float(bd.mean()), int(bd.max())],
                   "bp_schedule": args.schedule, "bp_sweeps_per_layer": sweeps_per_layer,
                   "sharding": ("none" if world == 1 else f"vertex row-strips over {world} ranks; NCCL: broadcast of level messages + Gram matrices, all-gather of the per-gate factorisation results (gate k solved on rank k mod {world})"),
                   "l2": "inputs larger than L2 (state %.2f GB)" % (sum(2 * int(np.prod([2] + [bd[e] for e, _ in g.incident[i]])) * 4
                                                                     for i in range(g.nv)) / 1e9),
                   "max_trunc_err": maxerr, "sz_center": zs, "step_ms": step_ms},
        "bp_sweep_ms": bp_sweep_ms,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
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
    ap.add_argument("--L", type=int, default=16)
    ap.add_argument("--chi", type=int, default=32)
    ap.add_argument("--prep", type=int, default=None, help="untimed layers before warm-up (default: 15 from the product state, 0 with --random-state)")
    ap.add_argument("--random-state", action="store_true", help="start from a synthetic random TNS with all bonds = chi (BASELINE config 5) instead of evolving the product state")
    ap.add_argument("--schedule", default="bipartite", choices=["bipartite", "forest"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--inplace", action="store_true", help="diagnostic: mutate the cache instead of the reference's functional copy")
    ap.add_argument("--no-sampler", action="store_true", help="diagnostic: do not poll nvidia-smi during the timed region")
    ap.add_argument("--cuda-profiler", action="store_true", help="cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)")
    ap.add_argument("--ref-budget", type=float, default=20.0)
    ap.add_argument("--ref-bp-sweeps", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
