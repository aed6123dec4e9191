#!/usr/bin/env python
"""One process per GPU (torchrun): parity and timing of the m-sharded transforms.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 \
        scripts/msharded_check.py --L 136 --B 2 [--bench]

Every rank also owns an unsharded plan of the same transform, so the expected
result is computed on the spot; rank 0 prints `MSHARDED OK` when every rank agrees
to 1e-12 relative L2, and one JSON line of timings with --bench.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=136)
    ap.add_argument("--B", type=float, default=2.0)
    ap.add_argument("--J_min", type=int, default=2)
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.pop("NCCL_DEBUG", None)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pxmcmc_b200 import device as D
    from pxmcmc_b200 import msharded as ms

    L, B, J = args.L, args.B, args.J_min
    ex = ms.ProcessGroupExchange()
    wp = ms.ShardedWaveletPlan(L, B, J, rank, world, exchange=ex)
    s0 = ms.ShardedShtPlan(L, 0, rank, world, exchange=ex)
    s2 = ms.ShardedShtPlan(L, 2, rank, world, exchange=ex)
    whole_w = D.WaveletPlan.get(L, B, J, 1)
    rng = np.random.default_rng(99)  # same stream on every rank
    coef = rng.standard_normal(wp.ncoefs) + 1j * rng.standard_normal(wp.ncoefs)
    pix = rng.standard_normal(L * (2 * L - 1)) + 1j * rng.standard_normal(L * (2 * L - 1))
    flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
    flm[:4] = 0
    worst = 0.0

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    checks = [("synthesis", coef, wp.coef_layout, wp.pix_layout), ("synthesis_adjoint", pix, wp.pix_layout, wp.coef_layout),
              ("analysis", pix, wp.pix_layout, wp.coef_layout), ("analysis_adjoint", coef, wp.coef_layout, wp.pix_layout)]
    for name, full_in, lin, lout in checks:
        expect = getattr(whole_w, name)(D.to_dev_c(full_in)).cpu().numpy()
        for _ in range(2):
            got = getattr(wp, name)(D.to_dev_c(lin.to_local(full_in))).cpu().numpy()
        e = rel(got, lout.to_local(expect)) if got.size else 0.0
        worst = max(worst, e)
        if rank == 0:
            print(f"wavelet {name}: rank0 rel-L2 {e:.2e}", flush=True)
    mask = ms.flm_owner_mask(L, rank, world)
    for spin, sp in ((0, s0), (2, s2)):
        whole = D.ShtPlan.get(L, spin, 1)
        for name, harm_in in (("inverse", True), ("forward_adjoint", True), ("forward", False), ("inverse_adjoint", False)):
            expect = getattr(whole, name)(D.to_dev_c(flm if harm_in else pix)).cpu().numpy()
            if harm_in:
                got = getattr(sp, name)(D.to_dev_c(np.where(mask, flm, 0))).cpu().numpy()
                e = rel(got, sp.pix_layout.to_local(expect)) if got.size else 0.0
            else:
                got = getattr(sp, name)(D.to_dev_c(sp.pix_layout.to_local(pix))).cpu().numpy()
                e = rel(got, np.where(mask, expect, 0))
            worst = max(worst, e)
            if rank == 0:
                print(f"sht spin {spin} {name}: rank0 rel-L2 {e:.2e}", flush=True)
    ok = wp.barrier_ok() and s0.barrier_ok() and s2.barrier_ok() and worst < 1e-12
    t = torch.tensor([0.0 if ok else 1.0, worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"worst rel-L2 over ranks {t[1].item():.2e}", flush=True)
        print("MSHARDED OK" if t[0].item() == 0.0 else "MSHARDED FAILED", flush=True)

    if args.bench:
        xc = D.to_dev_c(wp.coef_layout.to_local(coef))
        xp = D.to_dev_c(wp.pix_layout.to_local(pix))
        res = {"L": L, "B": B, "world": world}
        for name, x in (("synthesis", xc), ("synthesis_adjoint", xp)):
            for _ in range(5):
                getattr(wp, name)(x)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                getattr(wp, name)(x)
            e1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) / args.iters], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            res[f"sharded_{name}_ms"] = tt.item()
            xf = D.to_dev_c(coef if name == "synthesis" else pix)
            for _ in range(3):
                getattr(whole_w, name)(xf)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.iters):
                getattr(whole_w, name)(xf)
            e1.record()
            torch.cuda.synchronize()
            res[f"single_gpu_{name}_ms"] = e0.elapsed_time(e1) / args.iters
        if rank == 0:
            print(json.dumps(res), flush=True)
            if args.out:
                with open(args.out, "w") as f:
                    json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t[0].item() == 0.0 else 1)


if __name__ == "__main__":
    main()
