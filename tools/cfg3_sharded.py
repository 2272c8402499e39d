"""BASELINE cfg-3: GPT-XXL (1.4 B) c2i 384 px, global batch 256 sample-sharded over the ranks of one box (32 samples = 64 cache
rows per GPU at 8 ranks), 576 tokens, CFG 4.0, top-k 2000, one NCCL all_gather of the int32 code indices.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/cfg3_sharded.py

Rank 0 prints one JSON line: whole-job tokens/s (max over ranks of the CUDA-event time, barrier + synchronize on both sides)
and the fraction of the per-GPU HBM roofline of SURVEY.md 8d (25.4 k tok/s per GPU at the measured 6 544 GB/s)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
import torch.distributed as dist
from vlg.gpt import GPT_models
from vlg.dist import generate_sharded

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
GLOBAL_B, N = 32 * world, 576
torch.manual_seed(0)
with torch.device(dev):
    m = GPT_models["GPT-XXL"](block_size=576, cls_token_num=1, model_type="c2i").eval()
    torch.nn.init.normal_(m.output.weight, std=0.02)
m = m.to(dev, torch.bfloat16)
cond = torch.randint(0, 1000, (GLOBAL_B,), generator=torch.Generator().manual_seed(1234)).to(dev)
kw = dict(cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True)


def fence():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


out = generate_sharded(m, cond, N, global_seed=1, **kw)       # warm-up (graph capture, NCCL communicator)
fence()
steps = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    out = generate_sharded(m, cond, N, global_seed=2 + i, **kw)
e1.record()
fence()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
assert out.shape == (GLOBAL_B, N) and int(out.min()) >= 0 and int(out.max()) < 16384
same = torch.tensor([int(out.to(torch.int64).sum().item())], device=dev)       # every rank holds the same gathered indices
if world > 1:
    lo, hi = same.clone(), same.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert int(lo) == int(hi)
if rank == 0:
    v = GLOBAL_B * N / (float(ms) * 1e-3)
    print(json.dumps({"workload": f"cfg-3: GPT-XXL c2i 384px, global batch {GLOBAL_B} sample-sharded over {world} GPU(s) "
                                  "(32 samples = 64 rows per GPU), 576 tokens, bf16, CFG 4.0, top-k 2000, all_gather of indices",
                      "n_gpus": world, "ms_per_generate": round(float(ms), 1), "tokens_per_s": round(v, 1),
                      "tokens_per_s_per_gpu": round(v / world, 1), "hbm_roofline_tok_s_per_gpu": 25400,
                      "frac_of_hbm_roofline": round(v / world / 25400, 4), "steps": steps, "data": "synthetic"}))
if world > 1:
    dist.destroy_process_group()
