"""Worker of tests/test_bundle_sharded_gpu.py (one process per GPU, launched by torch.distributed.run):
sharded Bundle::Compute over NCCL against the single-GPU run of the same graph on every rank."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ptam_cg_b200 import synth  # noqa: E402
from ptam_cg_b200.capi import Bundle, nccl_unique_id, product_lib  # noqa: E402


def exchange_unique_id(lib, rank):
    uid = nccl_unique_id(lib) if rank == 0 else bytes(128)
    t = torch.tensor(list(uid), dtype=torch.uint8)
    dist.broadcast(t, 0)
    return bytes(t.tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cams", type=int, default=12)
    ap.add_argument("--points", type=int, default=800)
    ap.add_argument("--meas", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=21)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    lib = product_lib()
    uid = exchange_unique_id(lib, rank)
    g = synth.make_ba_graph(a.cams, a.points, a.meas, seed=a.seed)
    b = Bundle(lib, g["width"], g["height"], device=local)
    b.add_graph(g)
    b.init_shard(rank, world, uid)
    acc = b.Compute()
    pts, cams, outl, st = b.get_points(), b.get_cameras(), b.GetOutlierMeasurements(), b.stats()
    one = Bundle(lib, g["width"], g["height"], device=local)
    one.add_graph(g)
    acc1 = one.Compute()
    pts1, cams1, outl1, st1 = one.get_points(), one.get_cameras(), one.GetOutlierMeasurements(), one.stats()
    res = dict(rank=rank, accepted=(acc, acc1), trials=(st.lambda_trials, st1.lambda_trials),
               converged=(bool(b.Converged()), bool(one.Converged())), n_outliers=(len(outl), len(outl1)),
               outliers_equal=bool(np.array_equal(outl, outl1)),
               max_pt=float(np.abs(pts - pts1).max()), max_cam=float(np.abs(cams - cams1).max()),
               sigma=(st.sigma_squared, st1.sigma_squared))
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        print("SHARDED_BA_RESULT " + json.dumps(allres), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ok = (acc == acc1 and st.lambda_trials == st1.lambda_trials and res["outliers_equal"]
          and res["max_pt"] < 1e-6 and res["max_cam"] < 1e-6 and abs(st.sigma_squared - st1.sigma_squared) <= 1e-9 * st1.sigma_squared)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
