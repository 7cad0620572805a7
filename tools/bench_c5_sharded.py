"""Secondary benchmark (GPU box, torchrun): BASELINE.json configs[4] sharded by trajectory over the ranks --
65 536 reaction-diffusion grids of n = 2048 in total (strong scaling), RungeKutta45 rtol=1e-6 atol=1e-9, run([1.0]),
one CTA per trajectory, no collective in the data path; the final gather of the results is timed separately.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c5_sharded.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import torch.distributed as dist
import nbkode_b200 as nbk
from nbkode_b200 import distributed as nd
from nbkode_b200 import ensembles as onp  # input generators

B_total = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = 2048
rank, world = nd.rank_world()
local_rank = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
y0, p = onp.rd1d_ensemble(B_total, n=n, seed=3)
y0d = torch.from_numpy(nd.shard(y0, rank, world)).cuda()
pd = torch.from_numpy(nd.shard(p, rank, world)).cuda()
times, gather_ms, att, acc = [], 0.0, 0, 0
for it in range(3):
    s = nbk.RungeKutta45("rd1d", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    _, y = s.run([1.0])
    e1.record()
    full = nd.gather_batch(y if torch.is_tensor(y) else torch.from_numpy(y).cuda(), B_total)
    e2.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    gather_ms = e1.elapsed_time(e2)
    att = int((s._nacc.sum() + s._nrej.sum()).item())
    acc = int(s._nacc.sum().item())
    s.raise_for_status()
(ms, g_ms), (att_t, acc_t) = nd.reduce_max_sum([min(times[1:]), gather_ms], [att, acc], device="cuda")
if rank == 0:
    flop = 6 * 8 * n + 66 * n + 18
    print(json.dumps({"config": "C5 rd1d n=2048 sharded by trajectory", "n_gpus": world, "trajectories_total": B_total,
                      "ms_max_over_ranks": ms, "gather_ms": g_ms, "accepted_steps": acc_t, "attempts": att_t,
                      "accepted_steps_per_s": acc_t / ms * 1e3, "tflops_total": att_t * flop / ms * 1e3 / 1e12,
                      "result_shape": list(full.shape), "scaling": "strong"}))
if world > 1:
    dist.destroy_process_group()
