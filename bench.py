#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric (Mcell*stage/s of the inviscid per-RK-stage RHS path) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W]            the CUDA path (this repo)
    python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU implementation (oracle/_ref)

A "step" is one SSP-RK3 time step = 3 stages of {ghost fill, primitive recovery, x/y/z characteristic WENO sweeps, flux
divergence + stage update} plus the CFL dt.  Workload (config.workload): BASELINE.json configs[3], the 3-D multi-species
shock-bubble interaction (Inert-SBI, Emax = 9), WENO5-JS + LLF, 512^3 inner cells per GPU, weak-scaled in z over N GPUs with
a ghost-plane halo exchange per stage (NCCL send/recv) and a MAX all-reduce of the dt maxima per step.

value      = inner cells of all ranks * 3 * K / (device time of the K steps, max over ranks) / 1e6, state resident in HBM
N > 1      : the C++ slab stepper (csrc/xf_slab.cu: NCCL send/recv of the ghost planes overlapped with interior work, MAX all-reduce
             of dt) drives every rank; before anything is timed a small bitwise check of the decomposed run against the undecomposed
             block runs through the same stepper ("slab_check": "bitwise"), and after the weak-scaling measurement the jet
             (BASELINE configs[4], 1024x512x512 cut in z) is timed on the same ranks and on rank 0 alone ("strong": {...})
e2e        = the same through xf_step_host: every step uploads the AoS state from pinned host memory, runs the step, and
             downloads the AoS state (what the reference's per-step CopyToUbak does, src/XFLUIDS.cpp:174,644)
roofline   = the dominant kernel (the slowest directional sweep) against the FP64 FMA rate measured on this device
cpu_baseline = the unmodified reference (oracle/_ref, OpenMP host shim) on a bounded sample of the same case, rank 0, N = 1
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (settings json, sample grid, reference variant dir, E, NS)
    "sbi": dict(json="shock-bubble.json", grid=(512, 512, 512), ref="sbi_w5_fast", refcase="sbi", E=9, NS=5, cop=1,
                desc="shock-bubble.json Inert-SBI 3-D multi-species shock-bubble, WENO5-JS + LLF, inviscid, reactions off"),
    "jet": dict(json="expanded-jet.json", grid=(1024, 512, 512), ref="jet_w5_fast", refcase="jet", E=7, NS=3, cop=1,
                desc="expanded-jet.json 3-D under-expanded multi-component jet, WENO5-JS + LLF, inviscid"),
    # the 2-D single-component BASELINE configs (parity-test cases; measured with --no-cpu, no z decomposition)
    "riemann": dict(json="2d-riemann.json", grid=(4096, 4096, 0), ref="riemann_w5_fast", refcase="riemann", E=5, NS=1, cop=0,
                    desc="2d-riemann.json 2-D four-quadrant Riemann problem, single component, WENO5-JS + LLF"),
    "vortex": dict(json="2d-euler-vortex.json", grid=(1024, 1024, 0), ref="vortex_w5_fast", refcase="vortex", E=5, NS=1, cop=0,
                   desc="2d-euler-vortex.json 2-D isentropic Euler vortex, single component, WENO5-JS + LLF"),
}


# ---------------------------------------------------------------------------------------------------------------------
# algorithmic work model (BASELINE.md section 3 / SURVEY 8d); flop = FP64 add/mul/compare 1, FMA 2, div/sqrt/log 1
# ---------------------------------------------------------------------------------------------------------------------
def sweep_flops_per_face(E, NS, cop, weno=5):
    NC = NS - 1 if cop else 0
    B = 175 * NS + 40 * NC + 115 if cop else 4
    # per field: split + both reconstructions: WENO5-JS 212, WENO7-JS 399, WENO-CU6 358 (counted from WENO6s_schemes.hpp:5-78)
    return 32 + B + E * (34 * E + 2 * NC + {5: 212, 6: 358, 7: 399}[weno]) + 6 * E


def sweep_bytes_per_cell(E):
    return (E + 2) * 8 + E * 8  # read U, p, T once; write this direction's flux contribution once


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.p, self.f = index, None, None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])), mx.append(float(t[2])), pw.append(float(t[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), power_w_max=max(pw), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation of the path (oracle/_ref, else the oracle port)
# ---------------------------------------------------------------------------------------------------------------------
def run_cpu_reference(wl, grid, nsteps, threads):
    """Runs the unmodified reference (OpenMP host shim build) on `grid` for `nsteps`; returns (Mcell*stage/s, kind, seconds)."""
    w = WORKLOADS[wl]
    d = os.path.join(REPO, "oracle", "_ref", w["ref"])
    exe = os.path.join(d, "XFLUIDS")
    if os.path.exists(exe):
        out = os.path.join(d, "output")
        os.makedirs(os.path.join(out, "cal"), exist_ok=True)
        for f in os.listdir(out):
            if "CheckingPoint" in f or "AdaptiveRange" in f:
                os.remove(os.path.join(out, f))
        env = dict(os.environ, XF_NSTEPS=str(nsteps), OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="close")
        env.pop("XF_DUMP_DIR", None)
        r = subprocess.run([exe, "-run=%d,%d,%d,%d" % (grid[0], grid[1], grid[2], nsteps)], cwd=d, env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        m = re.search(r"ORACLE_TIMING steps=(\d+) seconds=([0-9.eE+-]+) mcell_stage_per_s=([0-9.eE+-]+)", r.stdout)
        if m and int(m.group(1)) == nsteps:
            return float(m.group(3)), "reference", float(m.group(2)), threads
        sys.stderr.write("bench: oracle/_ref run failed, falling back to the oracle port\n" + r.stdout[-500:] + "\n")
    # the oracle port (single thread): test infrastructure used here only as the timed CPU baseline
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import numpy as np
    import xfref
    from xfluids_b200 import host
    so = os.path.join(REPO, "oracle", "_build", "liboracle_fast.so")
    if not os.path.exists(so):
        subprocess.check_call(["bash", os.path.join(REPO, "oracle", "build_oracle.sh")])
    s = host.Setup(os.path.join(REPO, "settings", w["json"]), ["-run=%d,%d,%d" % tuple(grid), "-weno=5", "-alpha=LLF"])
    U, T = s.initial_condition()
    o = xfref.Oracle(w["refcase"], tuple(grid), weno=5, so=so)
    o.set_state(U, T)
    o.startup()
    t0 = time.time()
    n, _, _ = o.run(nsteps)
    sec = time.time() - t0
    return grid[0] * grid[1] * grid[2] * 3.0 * abs(n) / sec / 1e6, "port", sec, 1


def bind_to_gpu_numa(local):
    """N > 1: pin this rank to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffer is allocated and first
    touched, so that the e2e leg's 10 GB buffer lives on the NUMA node next to the GPU's PCIe root.  Best effort; returns a note."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "rank pinned to %d GPU-local CPUs" % len(cpus)
        return "no GPU-local CPUs in this cgroup"
    except Exception as e:  # noqa: BLE001
        return "not pinned (%s)" % type(e).__name__


def cpu_sample_grid(wl):
    # ~12 s of work for the 16 host threads of a pool box at 5 steps (x3 stages): large enough that the OpenMP loops are not dominated by
    # fork/join, small enough for the few-minutes budget; the reference cannot allocate 512^3 (int cellbytes, 157 GB)
    return (256, 128, 128)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    grid = cpu_sample_grid(wl)
    threads = os.cpu_count() or 1
    for _ in range(max(args.warmup, 0) and 1):
        run_cpu_reference(wl, grid, 1, threads)                     # one warm-up pass (page-in of the binary and the tables)
    val, kind, sec, used = run_cpu_reference(wl, grid, max(args.steps, 1), threads)
    w = WORKLOADS[wl]
    sample = "%s, %dx%dx%d inner cells, %d steps (x3 stages), one process, %d OpenMP threads" % (w["ref"] if kind == "reference" else "oracle port", grid[0], grid[1], grid[2], args.steps, used)
    line = {"impl": "reference", "metric": "cell-updates/sec (Mcell*stage/s)", "value": val, "unit": "Mcell*stage/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3 / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": w["desc"] + "; CPU sample " + "x".join(map(str, grid)) + " (the reference cannot allocate 512^3: int cellbytes, 157 GB)",
                       "grid_per_gpu": list(grid), "flush": "n/a"},
            "cpu_baseline": {"value": val, "unit": "Mcell*stage/s", "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Mcell*stage/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def slab_check(rank, world, local, dev, steps=3):
    """N > 1, before anything is timed: a small shock-bubble block (32 x 16 x 16 per rank, weak) advanced `steps` steps by the C++
    slab stepper on all ranks must equal, BIT FOR BIT, the undecomposed block advanced on rank 0's GPU alone (strict mode: no
    floating-point reduction crosses ranks except max).  Returns "bitwise" on every rank or raises."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from xfluids_b200 import capi, host
    from xfluids_b200.slab import SlabStepper
    js, grid, dom = os.path.join(REPO, "settings", "shock-bubble.json"), (32, 16, 16), (0.1, 0.05, 0.05)
    cli = ["-weno=5", "-alpha=LLF", "-pp=0"]
    mine = host.Setup(js, ["-run=%d,%d,%d" % grid, "-mpi=1,1,%d" % world, "-mpi-s=weak"] + cli, rank=rank, nranks=world)
    Um, Tm = mine.initial_condition()
    eng = capi.Engine(mine.block, mine.thermal, mine.scheme, device=local, keepalive=(mine,))
    st = SlabStepper(eng, mine.bc, rank, world, dev)
    eng.set_state(Um, Tm)
    st.startup()
    done, t, err = st.run(steps)
    assert (done, err) == (steps, 0), (done, err)
    Bz, zi, E = mine.block.Bwidth_Z, mine.block.Z_inner, mine.Emax
    plane = mine.block.Xmax * mine.block.Ymax * E
    slab = eng.download(eng.U).reshape(mine.block.Zmax, plane)[Bz:Bz + zi]
    st.close()
    eng.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (slab, t))
    ok = [1]
    if rank == 0:
        one = host.Setup(js, ["-run=%d,%d,%d" % (grid[0], grid[1], grid[2] * world), "-domain=%.17g,%.17g,%.17g" % (dom[0], dom[1], dom[2] * world)] + cli)
        U1, T1 = one.initial_condition()
        e1 = capi.Engine(one.block, one.thermal, one.scheme, device=local, keepalive=(one,))
        e1.set_state(U1, T1)
        e1.boundary(e1.U, one.bc)
        assert e1.update_states(e1.U) == 0
        d1, t1, err1 = e1.run(one.bc, steps)
        Uone = e1.download(e1.U).reshape(one.block.Zmax, plane)
        e1.close()
        ok = [int((d1, err1) == (steps, 0) and all(g[1] == t1 and np.array_equal(g[0], Uone[Bz + r * zi:Bz + (r + 1) * zi]) for r, g in enumerate(gathered)))]
    dist.broadcast_object_list(ok, src=0)
    if not ok[0]:
        raise SystemExit("bench.py: slab_check FAILED -- the decomposed run differs from the undecomposed block")
    return "bitwise"


def time_steps(run_steps, barrier, stream, dev, world, nwarm, nsteps):
    """W warm-up steps, then K steps between CUDA events on the launching stream, barrier + synchronize on both sides; ms = max over ranks."""
    import torch
    import torch.distributed as dist
    run_steps(nwarm)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_steps(nsteps)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def strong_leg(args, rank, world, local, dev, stream, barrier):
    """BASELINE configs[4]: the under-expanded jet, 1024 x 512 x 512, cut into `world` z-slabs (strong scaling), and the same box on
    rank 0's GPU alone for the efficiency denominator.  Returns the sub-record (rank 0) or None."""
    import numpy as np
    import torch
    from xfluids_b200 import capi, host
    from xfluids_b200.slab import SlabStepper
    w = WORKLOADS["jet"]
    grid = w["grid"]
    cli = ["-run=%d,%d,%d" % grid, "-weno=5", "-alpha=LLF", "-fp=%d" % args.fp, "-pp=0"]
    js = os.path.join(REPO, "settings", w["json"])
    cells = grid[0] * grid[1] * grid[2]
    nst, nwarm = max(3, min(args.steps, 5)), 3
    out = {}
    for leg in ("slabs", "one"):
        if leg == "one" and rank != 0:
            barrier()
            continue
        setup = host.Setup(js, cli + (["-mpi=1,1,%d" % world, "-mpi-s=strong"] if leg == "slabs" else []), rank=rank if leg == "slabs" else 0, nranks=world if leg == "slabs" else 1)
        U, T = setup.initial_condition()
        eng = capi.Engine(setup.block, setup.thermal, setup.scheme, device=local, keepalive=(setup,))
        eng.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            eng.set_state(U, T)
            del U, T
            if leg == "slabs":
                st = SlabStepper(eng, setup.bc, rank, world, dev)
                st.startup()
                ms = time_steps(st.steps, barrier, stream, dev, world, nwarm, nst)
                assert not st.any_error(), "numerical guard fired in the strong-scaling leg"
                st.close()
            else:
                eng.boundary(eng.U, setup.bc)
                assert eng.update_states(eng.U) == 0
                ms = time_steps(lambda n: eng.run(setup.bc, n), lambda: torch.cuda.synchronize(), stream, dev, 1, nwarm, nst)
                assert eng.error_flags()[:3] == [0, 0, 0]
        eng.close()
        out[leg] = cells * 3.0 * nst / (ms * 1e-3) / 1e6
        if leg == "one":
            barrier()
    if rank != 0:
        return None
    return {"workload": w["desc"] + "; %dx%dx%d inner cells cut into %d z-slabs" % (grid[0], grid[1], grid[2], world), "scaling": "strong", "n_gpus": world, "steps": nst,
            "value": out["slabs"], "unit": "Mcell*stage/s", "n1_value": out["one"], "efficiency_vs_n1": out["slabs"] / (world * out["one"]),
            "note": "n1_value: the same box on rank 0's GPU alone, same run (xf_run graph replay)"}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sbi", choices=sorted(WORKLOADS))
    ap.add_argument("--grid", default=None, help="nx,ny,nz inner cells (per GPU when weak, of the whole box when strong; default: the workload's BASELINE size)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="default: weak for sbi (z extent grows with N), strong for jet (fixed box cut in z)")
    ap.add_argument("--fp", type=int, default=0, help="0 strict (parity mode, default), 1 FMA contraction in the sweeps")
    ap.add_argument("--weno", type=int, default=5, help="5 WENO5-JS, 6 WENO-CU6, 7 WENO7-JS")
    ap.add_argument("--pp", type=int, default=0, help="1: positivity-preserving flux limiter on (fused into the sweep tails)")
    ap.add_argument("--visc", type=int, default=0, help="1: viscous + heat-conduction + species-diffusion wall fluxes on (the shipped shock-bubble preset: --weno 6 --pp 1 --alpha GLF --visc 1)")
    ap.add_argument("--alpha", default="LLF", choices=["ROE", "LLF", "GLF"], help="flux splitting (north_star: LLF)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--host-chunks", type=int, default=None, help="z-chunks of the overlapped upload in the e2e leg (xf_set_host_overlap; 0 = plain sequence; default: the library's 32)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling jet sub-record")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    # stdout carries exactly ONE line, the JSON record: whatever native libraries print on fd 1 (NCCL's version banner at communicator
    # creation) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if os.environ.get("WORLD_SIZE") and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        # torchrun pins OMP_NUM_THREADS=1; the host-side initial condition (OpenMP) gets this rank's share of the cores
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ["WORLD_SIZE"])))
    import numpy as np
    import torch
    import torch.distributed as dist
    from xfluids_b200 import capi, host
    from xfluids_b200.slab import SlabStepper

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- xfluids_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note = None
    if world > 1:
        numa_note = bind_to_gpu_numa(local)
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    w = WORKLOADS[args.workload]
    grid = tuple(int(x) for x in args.grid.split(",")) if args.grid else w["grid"]
    scaling = args.scaling or ("strong" if args.workload == "jet" else "weak")
    cli = ["-run=%d,%d,%d" % grid, "-weno=%d" % args.weno, "-alpha=" + args.alpha, "-fp=%d" % args.fp, "-pp=%d" % args.pp, "-visc=%d" % args.visc]
    if world > 1:
        cli += ["-mpi=1,1,%d" % world, "-mpi-s=%s" % scaling]
    setup = host.Setup(os.path.join(REPO, "settings", w["json"]), cli, rank=rank, nranks=world)
    E = setup.Emax
    ncells = setup.ncells
    inner = setup.block.X_inner * setup.block.Y_inner * setup.block.Z_inner
    L = capi.Lib.get()
    check = slab_check(rank, world, local, dev) if world > 1 else None

    # ---- initial condition written straight into pinned host memory (the e2e leg's host buffer) ----
    nbytes = ncells * E * 8
    hptr = L.dll.xf_host_alloc_pinned(nbytes)
    if not hptr:
        raise SystemExit("cannot pin %d bytes of host memory" % nbytes)
    hU = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_double)), shape=(ncells * E,))
    t0 = time.time()
    _, hT = setup.initial_condition(U=hU)
    t_ic = time.time() - t0

    eng = capi.Engine(setup.block, setup.thermal, setup.scheme, device=local, keepalive=(setup,))
    if args.visc:
        eng.set_transport(setup.transport, keepalive=(setup,))
    ws_gb = eng.L.dll.xf_field_doubles(eng.ctx) * 8 * 6 / 1e9
    if args.host_chunks is not None:
        L.check(L.dll.xf_set_host_overlap(eng.ctx, args.host_chunks))
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        eng.upload(eng.U, hU)
        eng.upload(eng.U1, hU)
        eng.set_scalar("T", hT)
        del hT
        stepper = SlabStepper(eng, setup.bc, rank, world, dev)
        stepper.startup()

        def run_steps(n):
            if world == 1:
                done, _, err = eng.run(setup.bc, n)          # CUDA-graph replay, device-resident dt
                assert done == n and err == 0, (done, err)
            else:
                stepper.steps(n)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        run_steps(max(args.warmup, 3))
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = eng.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = time_steps(run_steps, barrier, stream, dev, world, 0, args.steps)
        clocks = sampler.stop() if rank == 0 else None
        launches = eng.launches() - l0
        assert not stepper.any_error(), "numerical guard fired during the timed region"
        value = inner * world * 3.0 * args.steps / (ms * 1e-3) / 1e6

        # ---- end to end through host buffers: upload AoS U, one step, download AoS U; every step ----
        e2e = None
        ke = args.e2e_steps if args.e2e_steps is not None else min(args.steps, 5)
        if ke > 0 and world == 1:
            L.check(L.dll.xf_download_aos(eng.ctx, eng.U, hptr))   # host buffer = current device state (T warm start matches it)
            eng.step_host(hptr, setup.bc, 1)                 # warm-up (staging buffer allocation)
            barrier()
            e0.record(stream)
            for _ in range(ke):
                done, err = eng.step_host(hptr, setup.bc, 1)
                assert done == 1 and err == 0
            e1.record(stream)
            barrier()
            mse = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            mse = float(mse.item())
            e2e = {"value": inner * world * 3.0 * ke / (mse * 1e-3) / 1e6, "unit": "Mcell*stage/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                   "steps": ke, "ms_per_step": mse / ke, "api": "xf_step_host (pinned AoS U up in z-chunks overlapped with the plane-local part of stage 1, 1 step, AoS U down)" if args.host_chunks != 0 else "xf_step_host (pinned AoS U up, 1 step, AoS U down; no overlap)"}
        elif ke > 0:
            # N > 1: per rank, upload -> K_e steps through the slab stepper -> download, all inside the timed region
            L.check(L.dll.xf_download_aos(eng.ctx, eng.U, hptr))
            barrier()
            e0.record(stream)
            overlapped = True
            for _ in range(ke):
                applied, herr = stepper.step_host(hptr)      # chunked upload / download overlapped with stages 1 and 3, halo exchanges in between
                assert herr == 0, "numerical guard fired in the e2e leg"
                if not applied:
                    overlapped = False
                    eng.upload(eng.U, hU)
                    stepper.step()
                    L.check(L.dll.xf_download_aos(eng.ctx, eng.U, hptr))
            e1.record(stream)
            barrier()
            mse = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(mse, op=dist.ReduceOp.MAX)
            mse = float(mse.item())
            e2e = {"value": inner * world * 3.0 * ke / (mse * 1e-3) / 1e6, "unit": "Mcell*stage/s", "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
                   "steps": ke, "ms_per_step": mse / ke, "api": ("xf_host_begin / xf_host_stage1_finish / xf_host_stage3 per rank (PCIe copies in z-chunks under stages 1 and 3), halo over NCCL" if overlapped
                           else "xf_upload_aos + slab step (halo over NCCL) + xf_download_aos per rank")}

        if e2e is not None:
            # what bounds e2e: this rank's PCIe link, measured with the same pinned buffer and nothing else running (h2d first: the
            # d2h pass then restores nothing we need -- the buffer is not used again), next to the rates the e2e step achieved
            a_, b_ = C.c_double(), C.c_double()
            barrier()
            L.check(L.dll.xf_measure_pcie(eng.ctx, C.c_void_p(hptr), nbytes, C.byref(a_), C.byref(b_)))
            link = torch.tensor([a_.value, b_.value], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(link, op=dist.ReduceOp.MIN)
            sec = e2e["ms_per_step"] * 1e-3
            e2e.update(h2d_gbs_per_rank=nbytes / sec / 1e9, d2h_gbs_per_rank=nbytes / sec / 1e9,
                       link_h2d_gbs=float(link[0].item()), link_d2h_gbs=float(link[1].item()),
                       copy_floor_ms=(nbytes / float(link[0].item()) + nbytes / float(link[1].item())) / 1e6,
                       limiter="PCIe: the two copies alone take copy_floor_ms of the step's ms_per_step (link rates: slowest rank, all ranks copying at once only in the e2e leg itself); "
                               "ghost cells travel too (the reference's Ubak layout), %.1f %% of the bytes" % (100.0 * (1.0 - inner / float(ncells))))

        # ---- per-kernel device times of eager steps (CUDA events on the launching stream) ----
        prof = None
        if rank == 0 and world == 1 and args.profile_steps > 0:
            acc = {}
            for _ in range(args.profile_steps):
                p = eng.profile_step(setup.bc)
                for k, v in p.items():
                    acc[k] = acc.get(k, 0.0) + v / args.profile_steps
            prof = acc
    strong = None
    if world > 1 and not args.no_strong and args.workload == "sbi" and args.grid is None:
        # free the weak-scaling state first: the jet box on rank 0 alone needs most of the HBM
        stepper.close()
        eng.close()
        L.dll.xf_host_free_pinned(hptr)
        hptr = None
        strong = strong_leg(args, rank, world, local, dev, stream, barrier)
    roof = None
    if prof:
        dfma, copy = capi.measure_peaks(local)
        dims = [d for d, on in zip(("sweep_x", "sweep_y", "sweep_z"), (setup.block.DimX, setup.block.DimY, setup.block.DimZ)) if on]
        top = max(dims, key=lambda k: prof[k])
        ax = {"sweep_x": 0, "sweep_y": 1, "sweep_z": 2}[top]
        n_in = [setup.block.X_inner, setup.block.Y_inner, setup.block.Z_inner]
        faces = 1
        for a in range(3):
            faces *= n_in[a] + 1 if a == ax else n_in[a]
        fl = sweep_flops_per_face(E, setup.num_species, setup.cop, args.weno) * faces      # per launch
        t_launch = prof[top] / 3.0 * 1e-3                                                  # 3 launches (stages) per step
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
        # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of the same command (profiles/), if the
        # capture was taken on this workload and grid
        traffic, traffic_src, executed = None, None, None
        tpath = os.path.join(REPO, "profiles", "traffic.json")
        tj, key = {}, ""
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            key = "%s:%dx%dx%d:weno%d" % (args.workload, setup.block.X_inner, setup.block.Y_inner, setup.block.Z_inner, args.weno)
            if getattr(args, "visc", 0):
                key += ":visc"     # the sweeps carry the viscous wall fluxes in their tails: their own instruction counts
            if key in tj and top in tj[key]:
                traffic, traffic_src = tj[key][top], tj.get("source")
                executed = tj[key].get("executed")
        # the honest utilisation figure first: FP64 instructions the kernel EXECUTES per face (ncu, committed capture) / launch time, against
        # the FP64 issue rate of this device (one warp instruction per 2 cycles per SM sub-partition = half the DFMA probe's flop rate)
        inst_face = sum(executed["fp64_thread_inst_per_face"].values()) if executed and "fp64_thread_inst_per_face" in executed else None
        issue_peak = dfma / 2.0 if dfma else None                                          # T thread-instructions / s
        frac_issue = (inst_face * faces / t_launch / 1e12 / issue_peak) if (inst_face and issue_peak) else None
        roof = {"bound": "fp64", "kernel": "k_sweep<%s>" % top[-1],
                "frac_issue": frac_issue, "fp64_inst_per_face_executed": inst_face, "issue_peak_tinst_per_s": issue_peak,
                "achieved": fl / t_launch / 1e12, "peak": dfma, "unit": "TFLOP/s",
                "frac": fl / t_launch / 1e12 / dfma if dfma else None, "frac_vs_nominal_37": fl / t_launch / 1e12 / 37.0,
                "peaks_used": "frac_issue and frac: the DFMA rate measured live on this device (xf_measure_peaks; profiles/r02_peaks.json has the committed probe "
                              "with clocks); frac_vs_nominal_37: BASELINE.md's nominal 37 TFLOP/s.  `achieved` counts the reference's dense arithmetic (SURVEY 8d "
                              "flop model), frac_issue the instructions actually executed",
                "traffic": traffic, "traffic_source": traffic_src,
                # `achieved` counts the reference's dense E x E arithmetic (SURVEY 8d flop model); what the kernel EXECUTES (structural
                # zeros skipped, strict mode = no FMA contraction) and how busy the FP64 pipe is, from the committed ncu capture:
                "executed_ncu": executed,
                "peak_source": "FP64 FMA rate measured live by xf_measure_peaks on this device (MEASURED_PEAKS.json holds no FP64 figure; nominal 37)",
                "hbm_view": {"achieved_gbs": sweep_bytes_per_cell(E) * inner / t_launch / 1e9, "peak_gbs": peaks.get("hbm_gbs", 6650.0),
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650", "copy_gbs_live": copy},
                "ms_per_launch": t_launch * 1e3, "flops_per_face": sweep_flops_per_face(E, setup.num_species, setup.cop, args.weno),
                "step_breakdown_ms": prof}
        # the two HBM-side kernels against the measured copy peak: algorithmic bytes per launch / live launch time, and the
        # DRAM bytes ncu saw for the same launch
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ncells_all = setup.ncells
        alg = {"lu_rk": (3 * E + 2 * E + E) * 8 * inner,                                   # 3 flux fields + U, U1 + the written field
               "prim": ((E + (1 if setup.cop else 0)) + (11 + setup.num_species + max(setup.num_species - 1, 0) if setup.cop else 6) + (max(setup.num_species - 1, 0) if setup.cop and setup.ghost_species else 0)) * 8 * ncells_all}
        other = {}
        for kname in ("prim", "lu_rk"):
            tl = prof[kname] / 3.0 * 1e-3
            if tl <= 0.0:      # the marching sweeps carry the divergence and the update: no separate kernel
                continue
            other[kname] = {"bound": "hbm", "ms_per_launch": tl * 1e3, "achieved_gbs": alg[kname] / tl / 1e9, "peak_gbs": hbm_peak,
                            "frac": alg[kname] / tl / 1e9 / hbm_peak, "traffic": (tj.get(key, {}).get(kname) if os.path.exists(tpath) else None)}
        roof["other_kernels"] = other

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        g = cpu_sample_grid(args.workload)
        threads = os.cpu_count() or 1
        val, kind, sec, used = run_cpu_reference(args.workload, g, 5, threads)
        cpu = {"value": val, "unit": "Mcell*stage/s", "cores": used, "kind": kind,
               "sample": "%s: %dx%dx%d inner cells, 5 steps (15 stages), %.1f s, %d OpenMP threads" % (w["ref"] if kind == "reference" else "oracle port", g[0], g[1], g[2], sec, used)}

    if rank == 0:
        line = {"metric": "cell-updates/sec (Mcell*stage/s)", "value": value, "unit": "Mcell*stage/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["desc"], "grid_per_gpu": [setup.block.X_inner, setup.block.Y_inner, setup.block.Z_inner], "emax": E, "weno": args.weno, "positivity_preserving": bool(args.pp), "flux_splitting": args.alpha, "viscous": bool(args.visc),
                           "fp_mode": "strict (no FMA contraction; parity mode)" if args.fp == 0 else "fast (FMA contraction in sweeps/LU/RK)",
                           "decomposition": "z-slabs x%d, halo = 4 planes of U per face per stage (NCCL send/recv)" % world if world > 1 else "single block",
                           "flush": "working set %.1f GB per GPU >> 126 MB L2, no explicit flush" % ws_gb,
                           "ic_seconds_host": round(t_ic, 2), "host_numa": numa_note},
                "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e}
        if check:
            line["slab_check"] = check
        if strong:
            line["strong"] = strong
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if hptr:
        L.dll.xf_host_free_pinned(hptr)
        stepper.close()
        eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
