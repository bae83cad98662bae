#!/usr/bin/env python
"""bench.py -- DIFFormer propagation-layer throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Step      = one `full_attention_conv(q, k, v, 'simple')` forward (pass 1 reduce -> [all-reduce] ->
            pass 2 apply) over one batch of synthetic [N,H,D] fp32 node tensors.
Workload  = BASELINE configs[2]: N=132 534 (ogbn-proteins shape), H=4, D=64, fp32.  With G>1 ranks
            every rank holds 132 534 rows of a G*132 534-node graph (weak scaling) and the pass-1
            partials (67.6 KB) are all-reduced over NCCL between the passes.
value     = node-updates/s with Q,K,V resident in HBM; e2e = same through the public Python API with
            pinned HOST tensors (H2D of Q,K,V and D2H of the output inside the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_NODES, HEADS, DIM = 132534, 4, 64
METRIC = "DIFFormer-layer node-updates/sec (full_attention_conv 'simple', N=132534 H=4 D=64 fp32 per GPU)"
UNIT = "node-updates/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock + clock-event reasons with an `nvidia-smi -lms 20` side process while the
    timed region runs (a Python thread starves behind the launch loop's GIL)."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        import subprocess
        self.proc, self.t0, self.t1 = None, None, None
        try:
            ident = str(dev.index if dev.index is not None else 0)
            try:
                ident = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
            except Exception:
                pass
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ident, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.proc.stdout.readline()          # first sample: the tool is up
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def summary(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except Exception:
            return out
        clocks, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in text.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.t0 is not None and not (self.t0 - 0.02 <= ts <= self.t1 + 0.02):
                    continue
                clocks.append(int(float(f[1])))
                mx = int(float(f[2]))
                for nm, val in zip(names, f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        clocks.sort()
        out.update({"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": mx, "samples": len(clocks),
                    "reasons": sorted(reasons)})
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path, all host threads
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, budget_s, rows=None):
    from oracle import difformer_oracle as O
    ncpu = os.cpu_count() or 1
    q, k, v = O.synthetic_qkv(N_NODES, HEADS, DIM, seed=123)

    # 16 on the GPU box), so give the reference its best thread count: one calibration step each
    cands = sorted({c for c in (ncpu, 64, 32, 16, 8) if c <= ncpu}, reverse=True)
    best = (None, float("inf"))
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            O.simple_attention_reference_chain(q[:32768], k[:32768], v[:32768])
            t0 = time.perf_counter()
            O.simple_attention_reference_chain(q[:32768], k[:32768], v[:32768])
            dt = time.perf_counter() - t0
            if dt < best[1]:
                best = (c, dt)
    threads = best[0]
    torch.set_num_threads(threads)
    with torch.no_grad():
        if rows is None:
            t0 = time.perf_counter()
            O.simple_attention_reference_chain(q, k, v)
            t_full = time.perf_counter() - t0
            frac = min(1.0, budget_s / max(t_full * (steps + warmup), 1e-9))
            rows = max(4096, int(N_NODES * frac))
        rows = min(rows, N_NODES)
        qs, ks, vs = q[:rows].contiguous(), k[:rows].contiguous(), v[:rows].contiguous()
        for _ in range(warmup):
            O.simple_attention_reference_chain(qs, ks, vs)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.simple_attention_reference_chain(qs, ks, vs)
        dt = (time.perf_counter() - t0) / steps
    return rows / dt, dt, rows, threads


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    rate, dt, rows, threads = cpu_reference_rate(args.steps, args.warmup, budget_s=120.0)
    sample = (f"{rows} of {N_NODES} rows per step (cost is linear in rows), H={HEADS} D={DIM} fp32, "
              f"oracle transcription of the einsum chain difformer.py:18-39 on torch CPU, {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"full_attention_conv('simple') N={N_NODES} H={HEADS} D={DIM} fp32 (BASELINE configs[2])",
                       "rows_per_step": rows},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        group = dist.group.WORLD

    from difformer_b200 import ops
    from oracle import difformer_oracle as O
    if args.simple_impl:
        ops.set_simple_impl(args.simple_impl)

    n_total = float(N_NODES * world)
    q, k, v = (t.to(dev) for t in O.synthetic_qkv(N_NODES, HEADS, DIM, seed=123 + rank))
    T = N_NODES * HEADS * DIM * 4

    comm = None
    if group is not None and args.collective == "nvlink":
        from difformer_b200.sharded import RowShardComm
        comm = RowShardComm(group)
    plen = int(ops.lib.dif_simple_partials_len(HEADS, HEADS, DIM, DIM))

    def step():
        if comm is not None:          # pass 1 + all-reduce over peer-mapped NVLink memory in ONE kernel
            ex = comm.exchange(plen, dev)
            fused = ex.fused_reduce(q, k, v)
            if fused is not None:
                return ops.simple_apply(q, fused[0], n_total, HEADS, DIM, prepared=fused[1])
            partials = ex.allreduce(ops.simple_partials(q, k, v, out=ex.next_slot()))
            return ops.simple_apply(q, partials, n_total, HEADS, DIM)
        partials, prepared = ops.simple_partials(q, k, v, with_prepared=True)
        if group is not None:
            dist.all_reduce(partials, group=group)
            prepared = None       # the pass-2 operand image only matches the un-reduced partials
        return ops.simple_apply(q, partials, n_total, HEADS, DIM, prepared=prepared)

    def barrier():
        if group is not None:
            dist.barrier(group=group)
        torch.cuda.synchronize(dev)

    # ---- parity spot check on the exact bench inputs (rank 0, N=1 only: fp64 oracle intermediates)
    parity = None
    if world == 1:
        out = step()
        want = O.simple_partials(q.double().cpu(), k.double().cpu(), v.double().cpu())
        parity = {"out_rel_err": O.rel_err(out, O.simple_apply(q.double().cpu(), want))}
        flat = ops.simple_partials(q, k, v).double().cpu()
        parity["S_rel_err"] = O.rel_err(flat[:HEADS * DIM * DIM].reshape(HEADS, DIM, DIM), want["S"])

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    collective = args.collective
    if comm is not None:
        # watchdog (csrc/comm.cu): a rank whose kernel waited 30 s for a peer's flag reports it here; then every rank
        # switches to the NCCL all-reduce of the partials so that the run still produces a valid number
        bad = torch.tensor([1.0 if comm.exchange(plen, dev).timed_out() else 0.0], dtype=torch.float32, device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
        if float(bad.item()) > 0:
            comm, collective = None, "nccl"
            if rank == 0:
                print("bench: the NVLink exchange timed out waiting for a peer; falling back to NCCL", file=sys.stderr)
            for _ in range(3):
                step()
            barrier()
    sampler = ClockSampler(dev)
    barrier()
    sampler.begin()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = []          # per-kernel events of a subset of steps (pass 1 / pass 2 split)
    ev[0].record()
    for i in range(args.steps):
        if i % 16 == 0 and world == 1:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            partials, prepared = ops.simple_partials(q, k, v, with_prepared=True)
            e[1].record()
            ops.simple_apply(q, partials, n_total, HEADS, DIM, prepared=prepared)
            e[2].record()
            kev.append(e)
        else:
            step()
    ev[1].record()
    barrier()
    sampler.end()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms = float(t.item())
    value = N_NODES * world / (ms * 1e-3)

    # ---- end to end through the public API with pinned host buffers
    import difformer
    qh, kh, vh = (x.cpu().pin_memory() for x in (q, k, v))
    oh = torch.empty((N_NODES, HEADS, DIM), dtype=torch.float32).pin_memory()
    rs = None
    if group is not None:
        from difformer_b200.sharded import RowShardedAttention
        rs = RowShardedAttention(int(n_total), group, nvlink=(collective == "nvlink"))

    # Double-buffered, three streams: the upload of step i+1 (copy engine, H2D) overlaps the kernels of step i and the
    # download of step i-1 (second copy engine, D2H).  Every step still uploads its own Q, K, V from pinned host
    # memory and downloads its own result; PCIe is full duplex, so the steady state is bound by the larger of the two.
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    dbuf = [tuple(torch.empty_like(x) for x in (q, k, v)) for _ in range(2)]
    ohs = [oh, torch.empty_like(oh).pin_memory()]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    live = [None, None]                       # keeps a step's device result alive until its download has been queued twice over

    def e2e_step(i):
        b = i & 1
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_cmp[b])        # the kernels of step i-2 have finished reading this input buffer
            for dst, src in zip(dbuf[b], (qh, kh, vh)):
                dst.copy_(src, non_blocking=True)
            ev_in[b].record(s_in)
        cur.wait_event(ev_in[b])
        with torch.no_grad():
            o = rs(*dbuf[b]) if rs is not None else difformer.full_attention_conv(*dbuf[b], "simple")
        ev_cmp[b].record(cur)
        o.record_stream(s_out)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[b])
            ohs[b].copy_(o, non_blocking=True)
            ev_out[b].record(s_out)
        live[b] = o

    def e2e_drain():
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev_out[0])
        cur.wait_event(ev_out[1])

    e2e_steps = max(3, min(args.steps, 20))
    for i in range(4):
        e2e_step(i)
    e2e_drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        e2e_step(i)
    e2e_drain()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    e2e_ms = float(t.item())

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg_bytes = 4 * T                      # read Q,K,V once + write out once
        achieved = alg_bytes / (ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "simple op = pass 1 (reduce, cross-CTA sum fused) + pass 2 (apply): one launch sequence per step",
                "algorithmic_bytes_per_step": alg_bytes}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tp):
            try:
                roof["traffic"] = json.load(open(tp)).get("simple_step_dram_bytes")
            except Exception:
                pass
        if kev:
            r_ms = sum(e[0].elapsed_time(e[1]) for e in kev) / len(kev)
            a_ms = sum(e[1].elapsed_time(e[2]) for e in kev) / len(kev)
            roof["passes"] = [
                {"name": "pass1 reduce (+fused finalize)", "ms": r_ms, "moved_bytes": 3 * T, "gbs": 3 * T / (r_ms * 1e-3) / 1e9},
                {"name": "pass2 apply", "ms": a_ms, "moved_bytes": 2 * T, "gbs": 2 * T / (a_ms * 1e-3) / 1e9}]
        torch_gpu = None
        if world == 1:
            # the reference's own op chain (einsums + materialised broadcasts) on the same B200, CUDA tensors
            with torch.no_grad():
                for _ in range(5):
                    O.simple_attention_reference_chain(q, k, v)
                torch.cuda.synchronize(dev)
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(20):
                    O.simple_attention_reference_chain(q, k, v)
                g1.record()
                torch.cuda.synchronize(dev)
            tg = g0.elapsed_time(g1) / 20
            torch_gpu = {"value": N_NODES / (tg * 1e-3), "unit": UNIT, "ms_per_step": tg,
                         "what": "reference einsum chain (difformer.py:18-39 transcription) in PyTorch eager on the same GPU"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, dt, rows, threads = cpu_reference_rate(steps=8, warmup=2, budget_s=20.0)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{rows} of {N_NODES} rows x 8 steps, oracle transcription of the einsum chain difformer.py:18-39, torch CPU fp32"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"full_attention_conv('simple') N={N_NODES} H={HEADS} D={DIM} fp32 per GPU (BASELINE configs[2])",
                           "rows_per_gpu": N_NODES, "global_rows": int(n_total),
                           "parallelism": "single GPU" if world == 1 else (
                               f"row-shard x{world}, one all-reduce of 16898 fp32 per step: " +
                               ("fused into the pass-1 kernel tail over peer-mapped NVLink memory (no NCCL call)" if collective == "nvlink" else "NCCL")),
                           "l2": "inputs 407 MB + output 136 MB per step exceed the 126 MB L2; no flush between steps",
                           "simple_impl": args.simple_impl or "auto"},
                "roofline": roof, "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu,
                "e2e": {"value": N_NODES * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": 3 * T, "d2h_bytes_per_step": T, "steps": e2e_steps,
                        "api": "difformer.full_attention_conv(q, k, v, 'simple') on pinned host tensors; double-buffered (upload of step i+1 overlaps download of step i-1)"},
                # tcgen05 path: reduce (cross-CTA sum fused in) + apply; generic path: reduce + finalize + apply
                "gpu_launches": ((3 if (args.simple_impl == "generic" or os.environ.get("DIF_TC_P1_TMA") == "0") else 2)
) * args.steps, "clocks": sampler.summary(), "parity": parity}
        print(json.dumps(line), flush=True)
    if group is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--simple-impl", default=None, choices=[None, "auto", "generic", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="nvlink", choices=["nvlink", "nccl"], help="multi-GPU all-reduce of the partials")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
