"""BASELINE config 5: contact-driven transport on the C2 film -- monte_carlo::init, then step / save_metrics /
repopulate_contacts iterations -- on one GPU or, under torchrun, sharded over several (contact populations split per
rank, one NCCL all-reduce of the integer bins per engine call).  Prints population, hop throughput and the mean net
crossings per interface.

    python tools/run_c5.py [c1_pop] [iterations] [steps_per_call]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_c5.py ...
"""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from cnt_film_monte_carlo_b200.parallel import ShardedContacts
from bench import mc_block, DT

c1 = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000      # whole-job contact population
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
per_call = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.current_stream(dev)
sc = ShardedContacts(c1, 0, 1, rank, world)
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
e = Engine(mc_block(1), device=local_rank, stream=stream.cuda_stream)
e.set_mesh(pos, ori)
if "CNTMC_C5_TOP" in os.environ:
    e.set_option("top_entries", int(os.environ["CNTMC_C5_TOP"]))   # 0: row search only
sc.configure(e)
t0 = time.time(); e.init(sc.c1_pop, sc.c2_pop, seed=sc.seed, capacity=int(7 * sc.c1_pop)); t_init = time.time() - t0
n_seg = e.number_of_segments()
bins = torch.zeros((per_call, 2 * n_seg - 1), dtype=torch.int64, device=dev)


def call():
    e.step_dev(DT, per_call, bins.data_ptr())
    return sc.bins(bins)                                 # NCCL all-reduce in place (a few kilobytes)


call(); e.sync()                                          # warm-up
P0 = e.number_of_particles(); h0 = e.hops()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); ev0.record(stream)
pops = []
for _ in range(iters // per_call):
    pops.append(call().clone())
ev1.record(stream)
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); e.sync()
ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
cnt = torch.tensor([float(e.hops() - h0), float(e.number_of_particles())], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(cnt)
if rank == 0:
    allb = torch.cat(pops).cpu().numpy()
    n_it = len(allb)
    print(json.dumps({"config": "C5: C2 film, contacts, c1_pop %d over %d GPU(s)" % (c1, world), "n_gpus": world, "excitons_now": int(cnt[1].item()),
                      "init_s": round(t_init, 2), "iterations": n_it, "device_ms_per_iteration": float(ms.item()) / n_it,
                      "hops_per_s": float(cnt[0].item()) / (float(ms.item()) * 1e-3),
                      "excitons_counted_per_iteration": float(allb[:, :n_seg].sum(1).mean()),
                      "mean_net_crossings": [float(x) for x in allb[:, n_seg:].mean(0)]}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
