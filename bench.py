"""Benchmark of the CFR hot path (BASELINE.json metric: CFR+ iterations/s, beside the CPU path on the same box).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload leduc_b5|leduc_pot|leduc_b3]

A "step" = one full CFR+ iteration (both seats: value/regret sweep + reach/average sweep each) on the whole public
tree, with the exact best-response evaluation of the current AND the average strategy every `--eval-every`
iterations (BASELINE.json config 2: "exact BR every 20 iters"; evaluations fall inside the timed region).
Default workload: DiscretizedNLLeduc with bet_sets.B_5, stack 20000 (873 586 nodes, 304 678 decision nodes,
sum of actions 860 103, range 6) - the largest Leduc tree of SURVEY.md §6/§8(d), whose working set (~190 MB of node
vectors + tables) exceeds the 126 MB L2.  The tree is deterministic: there is no dataset and no seed.

N > 1 (torchrun): the reference's `starting_stack_sizes` axis (`_CFRBase.py:44-69`: one independent tree per stack
size, results averaged) is spread over the ranks - rank r solves stack 20000 + 1000 r - with no data-path
collective ("weak" scaling); only the timing reduction uses NCCL.

Rank 0 prints ONE JSON line.  `--impl reference` times the CPU restatement of the reference's path (the C oracle,
pinned bit-for-bit to the reference; /root/reference itself does not exist on the GPU box) on the same workload with
all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    "leduc_b5": ("DiscretizedNLLeduc", "B_5"),
    "leduc_b3": ("DiscretizedNLLeduc", "B_3"),
    "leduc_pot": ("DiscretizedNLLeduc", "POT_ONLY"),
}


def make_tree(workload, stack):
    from pokerrl_b200.game import bet_sets, games
    from pokerrl_b200.game.flat_tree import FlatTree
    cn, bs = WORKLOADS[workload]
    g = getattr(games, cn)
    args = g.ARGS_CLS(n_seats=2, starting_stack_sizes_list=[stack, stack],
                      bet_sizes_list_as_frac_of_pot=list(getattr(bet_sets, bs)))
    return g, FlatTree(g, args)


def tree_stats(ft):
    import numpy as np
    dec = (ft.kind <= 1) & (ft.first_child >= 0)
    return dict(nodes=int(ft.n_nodes), decision=int(dec.sum()), sum_actions=int(ft.n_slots),
                terminal=int((ft.kind >= 3).sum()), levels=int(ft.n_levels), range=int(ft.R),
                sum_actions_p=[int(ft.n_children[dec & (ft.kind == p)].sum()) for p in (0, 1)],
                decision_p=[int((dec & (ft.kind == p)).sum()) for p in (0, 1)],
                nonterminal=int(((ft.kind <= 2) & (ft.first_child >= 0)).sum()),
                max_level_nodes=int(np.diff(ft.level_start).max()))


def algorithmic_bytes(st):
    """Minimum bytes the two sweeps of ONE seat must move in the level-synchronous design:
    value sweep: write ev[p] (N rows) + read every child's ev[p] once (N-1) + opponent reach at terminals (T) +
                 regret read/write and strategy read/write at the seat's decision nodes (4 rows per action) +
                 structure (kind 1 B, first_child 4 B, n_children 4 B per node; pot/board/acted_last 9 B per
                 terminal; first slot 4 B per decision node of the seat)
    reach sweep: write reach[p] (N) + read each non-terminal parent row once (NT) + strategy read and average
                 read/write at the seat's nodes (3 rows per action) + structure (parent 4 B, parent kind 1 B,
                 slot 4 B, board 4 B per node)
    Rows are range*4 bytes.  Returned per seat-averaged half-iteration: (value_bytes, reach_bytes)."""
    row = st["range"] * 4
    N, T, NT = st["nodes"], st["terminal"], st["nonterminal"]
    sa = sum(st["sum_actions_p"]) / 2.0
    dp = sum(st["decision_p"]) / 2.0
    value = row * (N + (N - 1) + T + 4 * sa) + 9 * N + 9 * T + 4 * dp
    reach = row * (N + NT + 3 * sa) + 13 * N
    return value, reach


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.path = None, "/tmp/prl_clocks_%d.csv" % os.getpid()
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax = float(c[2])
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_cpu(ft, n_iters, eval_every, threads):
    """C-oracle CFR+ on the host: returns seconds per iteration (including the evaluation cadence)."""
    import cfr_c
    s = cfr_c.OracleCSolver(ft, "CFRPlus", avg_f64=False, n_threads=threads)
    s.iteration(1)  # warm the caches / page in
    t0 = time.perf_counter()
    for i in range(n_iters):
        s.iteration(1)
        if (i + 1) % eval_every == 0:
            s.exploitability_current()
            s.exploitability_average()
    dt = time.perf_counter() - t0
    return dt / n_iters, s.n_threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="leduc_b5", choices=list(WORKLOADS))
    ap.add_argument("--eval-every", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(3, a.warmup)
    K = a.steps
    stack = 20000 + 1000 * rank
    cfg = {"workload": "%s CFR+ delay 0, %s, stack 20000%s, exact BR (current+average) every %d iterations" % (
        WORKLOADS[a.workload][0], "bet_sets." + WORKLOADS[a.workload][1],
        " + 1000*rank (one tree per rank)" if world > 1 else "", a.eval_every)}

    # ------------------------------------------------------------------ reference arm (CPU restatement)
    if a.impl == "reference":
        if rank != 0:
            return
        g, ft = make_tree(a.workload, 20000)
        st = tree_stats(ft)
        ncpu = os.cpu_count() or 1
        K = min(K, 200)
        sec, threads = run_cpu(ft, K, a.eval_every, min(ncpu, 16))
        cfg.update(tree=st)
        v = 1.0 / sec
        print(json.dumps({
            "impl": "reference", "metric": "CFR+ iterations/s", "value": v, "unit": "iterations/s", "n_gpus": a.gpus,
            "steps": K, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (deterministic game tree, no dataset)",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": threads, "kind": "port",
                             "sample": "%d full CFR+ iterations of the same tree by oracle/cfr_oracle.c (OpenMP)" % K},
            "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ B200 arm
    import ctypes as C

    import torch
    import torch.distributed as dist
    from pokerrl_b200 import _native as nat
    from pokerrl_b200.solver import CFRSolver, _stream

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t0 = time.perf_counter()
    g, ft = make_tree(a.workload, stack)
    t_build = time.perf_counter() - t0
    st = tree_stats(ft)
    t0 = time.perf_counter()
    s = CFRSolver(ft, "CFRPlus", delay=0, avg_f64=False)
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t0
    tree_bytes = sum(getattr(s.dtree, k).numel() * getattr(s.dtree, k).element_size()
                     for k in ("t_parent", "t_first_child", "t_n_children", "t_slot", "t_kind", "t_acted_last",
                               "t_pot", "t_board"))

    def step(i):
        s.iteration(1)
        if (i + 1) % a.eval_every == 0:
            return s.exploitability_current(), s.exploitability_average()
        return None

    def steps(i0, n):
        """n steps starting at step index i0; iterations between two evaluations share one persistent launch"""
        out, i = [], i0
        while i < i0 + n:
            m = min(a.eval_every - (i % a.eval_every), i0 + n - i)
            s.iteration(m)
            i += m
            if i % a.eval_every == 0:
                out.append((i, s.exploitability_current(), s.exploitability_average()))
        return out

    for i in range(W):
        step(i)
    s.reset()  # timed run starts from iteration 0 so that the exploitability trace is the reference's
    for i in range(W):
        step(i)
    s.reset()
    torch.cuda.synchronize()

    # --- timed region: K steps, device-timed with CUDA events on the launching stream
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        time.sleep(0.5)  # let nvidia-smi start sampling; the GPU is kept busy by an untimed step stream meanwhile
        for i in range(W):
            s.iteration(1)
        s.reset()
        torch.cuda.synchronize()
    launches0 = nat.lib().prl_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace = []
    wall0 = time.perf_counter()
    ev0.record()
    trace = steps(0, K)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    dev_ms = ev0.elapsed_time(ev1)
    launches = nat.lib().prl_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())

    # --- e2e: the user-facing call (CFRPlus façade: iteration + logging through ChiefBase, results read on host)
    from pokerrl_b200.cfr.CFRPlus import CFRPlus
    from pokerrl_b200.game import bet_sets, games
    from pokerrl_b200.rl.base_cls.workers.ChiefBase import ChiefBase
    del s
    torch.cuda.empty_cache()
    cn, bs = WORKLOADS[a.workload]
    chief = ChiefBase(t_prof=None)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        cfr = CFRPlus(name="bench", chief_handle=chief, game_cls=getattr(games, cn),
                      agent_bet_set=list(getattr(bet_sets, bs)), starting_stack_sizes=[stack], delay=0,
                      eval_every=a.eval_every)
    for _ in range(W):
        cfr.iteration()
    cfr.reset()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = time.perf_counter()
    for _ in range(K):
        cfr.iteration()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0
    t2 = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_s = float(t2.item())
    n_evals = K // a.eval_every
    d2h_per_step = 2 * 8 * n_evals / K  # two float32[2] exploitability read-backs per evaluation

    # --- roofline of the dominant kernel: the persistent cooperative kernel that runs whole CFR+ iterations
    # (cfr_iterations_kernel, one launch per `eval_every` iterations), timed live with CUDA events on its stream
    s = cfr.solvers[0]
    it_ms = []
    for rep in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.iteration(a.eval_every)
        e1.record()
        torch.cuda.synchronize()
        if rep >= 2:
            it_ms.append(e0.elapsed_time(e1))
    vb, rb = algorithmic_bytes(st)
    bytes_per_launch = a.eval_every * 2 * (vb + rb)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    l_ms = statistics.mean(it_ms)
    achieved = bytes_per_launch / (l_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "cfr_iterations_kernel<6,2> (persistent cooperative kernel: %d CFR+ iterations = "
                "%d level steps with grid barriers per launch)" % (a.eval_every, a.eval_every * 2 * (2 * st["levels"] - 1)),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": l_ms,
                "algorithmic_bytes_per_iteration": 2 * (vb + rb), "traffic": None,
                "note": "latency/occupancy-bound, not HBM-bound: R = 6 rows, 27 dependent level steps per seat"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {
        "metric": "CFR+ iterations/s", "value": world * K / (max_ms * 1e-3), "unit": "iterations/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (deterministic game tree, no dataset)",
        "config": dict(cfg, tree=st, l2="working set > L2: node vectors %d MB + tables %d MB + structure %d MB; no explicit flush" % (
            4 * st["nodes"] * st["range"] * 4 // 2 ** 20, 3 * st["sum_actions"] * st["range"] * 4 // 2 ** 20,
            tree_bytes // 2 ** 20), parallelism="one tree per rank (stack-size axis), no collective"),
        "clocks": clocks,
        "e2e": {"value": world * K / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": d2h_per_step,
                "note": "CFRPlus.iteration() façade incl. ChiefBase logging; CFR has no per-step host input - the "
                        "one-off tree upload is reported under setup"},
        "setup": {"tree_build_s": t_build, "tree_upload_s": t_upload, "tree_h2d_bytes": tree_bytes},
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall * 1e3 / K,
        "exploitability_trace_mbb_per_g": trace[-3:],
        "roofline": roofline,
    }
    if world == 1 and not a.no_cpu_baseline:
        ncpu = os.cpu_count() or 1
        per_iter_guess = 0.25 * st["nodes"] / 873586.0
        n = max(2, min(K, int(15.0 / max(per_iter_guess, 1e-4))))
        n = (n // a.eval_every) * a.eval_every or n
        _, ft0 = (g, ft)
        sec, threads = run_cpu(ft0, n, a.eval_every, min(ncpu, 16))
        out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "iterations/s", "cores": threads, "kind": "port",
                               "sample": "%d full CFR+ iterations (same tree, same BR cadence) by oracle/cfr_oracle.c "
                                         "with OpenMP; the reference's own Python path is ~400x slower per node "
                                         "(BASELINE.md: 0.448 s/iter on the 1 096-node tree)" % n}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
