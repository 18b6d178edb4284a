"""Sharded Bundle::Compute on C4 alone (no tracker part): launched with torch.distributed.run on N GPUs, prints the
lambda-trials/s of the sharded run, its phases, and the same graph on rank 0's GPU alone."""
import json, os, sys
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ptam_cg_b200 import capi, synth
from ptam_cg_b200.bench_ba import CONFIGS, bench_ba

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prod = capi.product_lib()
    g4 = synth.make_ba_graph(**CONFIGS["C4"])
    uid = torch.tensor(list(capi.nccl_unique_id(prod) if rank == 0 else bytes(capi.NCCL_UNIQUE_ID_BYTES)), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    comm = capi.nccl_comm_create(prod, local, rank, world, bytes(uid.cpu().tolist()))
    r = bench_ba(prod, local, "C4", reps=3, shard=(rank, world, comm), graph=g4)
    prod.fn("nccl_comm_destroy")(comm)
    r.pop("_result")
    allp = [None] * world
    dist.all_gather_object(allp, r["phases_ms_per_call"])
    if rank == 0:
        one = bench_ba(prod, local, "C4", reps=2, graph=g4)
        one.pop("_result")
        print(json.dumps({"n_gpus": world, "sharded": {k: r[k] for k in ("value", "compute_ms", "lambda_trials", "phases_ms_per_call")},
                          "single": {k: one[k] for k in ("value", "compute_ms", "lambda_trials", "phases_ms_per_call")},
                          "per_rank_phases_ms_per_call": allp}), flush=True)
    dist.barrier()
    dist.destroy_process_group()

main()
