#!/usr/bin/env python
"""BASELINE config 5 as SURVEY.md 8(d) states it: chain-count x GPU-count sweep on 2-D Rosenbrock,
N in {2^10 ... 2^24} total chains over R in {1, 2, 4, 8} GPUs, nburn 500 + nsamp 1000, thin 10, PLOCAL 0.9,
pool M = min(N, 256), Murray exchange included.  ONE torchrun launch covers every (N, R) point: the first R
ranks of the job form the group of a point, the others idle at a barrier (process start-up would otherwise
dominate the GPU-minutes).

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_c5.py [--modes summix,reference]
  python tools/sweep_c5.py                       # one GPU: the R = 1 column

Metric per point: total chain-steps of the whole device loop (burn-in incl. tuning, main loop incl. exchange,
history ring writes) / CUDA-event time on the launching stream, max over the point's ranks.
Writes gpurun_out/sweep_c5_<world>gpu.json and .md.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                                        # noqa: E402  (Job: engine + exchange wiring of one rank)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="summix,reference")
    ap.add_argument("--min-exp", type=int, default=10)
    ap.add_argument("--max-exp", type=int, default=24)
    ap.add_argument("--ref-max-exp", type=int, default=22, help="largest N for the reference remote mode at M = 256 (its remote step costs ~M candidates)")
    ap.add_argument("--pool", type=int, default=256)
    ap.add_argument("--nsamp", type=int, default=1000)
    ap.add_argument("--workload", default="rosen2d")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    Rs = [r for r in (1, 2, 4, 8) if r <= world]
    groups = {world: None}
    for r in Rs:
        if 1 < r < world:
            groups[r] = dist.new_group(ranks=list(range(r)))         # every rank calls new_group for every group
    args = argparse.Namespace(workload=a.workload, pl=0.9, sync=10, thin=10, coin_group=0, lag=1, exchange="p2p")
    ctx = {"args": args, "W": dict(bench.WORKLOADS[a.workload]), "rank": rank, "world": world, "local": local, "dev": dev,
           "stream": stream, "dist": dist, "groups": groups}
    nburn, nsamp = 500, a.nsamp
    rows = []
    for mode in a.modes.split(","):
        rm = bench.RMODE[mode]
        for e in range(a.min_exp, a.max_exp + 1, 2):
            N = 1 << e
            if mode == "reference" and e > a.ref_max_exp:
                continue
            for R in Rs:
                Cg = N // R
                if Cg < 32 or Cg % 32:
                    continue
                M = min(N, a.pool)
                if world > 1:
                    dist.barrier()
                job = bench.Job(ctx, Cg, rm, M, ring=8, ranks=R)
                if job.active:
                    job.e.set_state(job.pinit())
                    torch.cuda.synchronize(); job.barrier()
                    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
                    ev0.record(stream)
                    job.burn(nburn)
                    job.e.sample_begin(nsamp)
                    for _ in range(nsamp // args.sync):
                        job.window()
                    ev1.record(stream)
                    torch.cuda.synchronize()
                    ms = ev0.elapsed_time(ev1)
                    if R > 1:
                        t = torch.tensor([ms], dtype=torch.float64, device=dev)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=groups[R])
                        ms = float(t.item())
                    st = job.e.stats()
                    if rank == 0:
                        rows.append({"mode": mode, "N": N, "log2N": e, "R": R, "M": M, "value": N * (nburn + nsamp) / (ms * 1e-3), "ms": ms,
                                     "remote_iterations_mean": st["remote_iterations"] / max(1, st["remote_steps"]),
                                     "accept_rate": st["accepted"] / max(1, st["tried"]),
                                     "exchange_wait_ms": st["exchange_wait_ns"] * 1e-6})
                        print(json.dumps(rows[-1]), flush=True)
                    job.close()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        base = os.path.join(ROOT, "gpurun_out", "sweep_c5_%dgpu" % world)
        json.dump(rows, open(base + ".json", "w"), indent=1)
        md = ["# Config 5 (SURVEY.md 8d): Rosenbrock d=2, N total chains over R GPUs, nburn 500 + nsamp %d, thin 10, PLOCAL 0.9, pool M = min(N, %d), pool lag 1, p2p exchange" % (nsamp, a.pool), ""]
        for mode in a.modes.split(","):
            md += ["## remote mode: %s" % mode, "", "| N | " + " | ".join("R=%d chain-steps/s" % r for r in Rs) + " | " + " | ".join("eff R=%d" % r for r in Rs[1:]) + " | iterations / remote step |", "|---|" + "---|" * (2 * len(Rs))]
            for e in range(a.min_exp, a.max_exp + 1, 2):
                pts = {r["R"]: r for r in rows if r["mode"] == mode and r["log2N"] == e}
                if not pts:
                    continue
                v1 = pts.get(1, {}).get("value")
                cells = ["%.3g" % pts[r]["value"] if r in pts else "-" for r in Rs]
                effs = ["%.2f" % (pts[r]["value"] / v1) if (r in pts and v1) else "-" for r in Rs[1:]]
                md.append("| 2^%d | %s | %s | %.1f |" % (e, " | ".join(cells), " | ".join(effs), list(pts.values())[0]["remote_iterations_mean"]))
            md.append("")
        md.append("eff R=k: the R=k value over the R=1 value at the SAME total N (strong scaling of a fixed job; 1.0 = no gain, k = linear).")
        open(base + ".md", "w").write("\n".join(md) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
