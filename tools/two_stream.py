"""Experiment: split the cfg-2 batch into two independent half-batches on two streams (latency bubbles of one half's
GEMM chain overlap the other half's attention streaming)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg.gpt import GPT_models, Engine

torch.manual_seed(0)
model = GPT_models["GPT-L"](block_size=576, cls_token_num=1, model_type="c2i").eval()
torch.nn.init.normal_(model.output.weight, std=0.02)
model = model.to("cuda", torch.bfloat16)
cond = torch.randint(0, 1000, (64,), generator=torch.Generator().manual_seed(1234)).cuda()
N = 576
kw = (4.0, -1, 1.0, 2000, 1.0, True)

def run(parts):
    engines = [Engine(model, 2 * c.shape[0], 1 + N) for c in parts]
    streams = [torch.cuda.Stream() for _ in parts]
    def once(seed):
        outs = []
        for e, s, c in zip(engines, streams, parts):
            with torch.cuda.stream(s):
                outs.append(e.generate(c, N, None, *kw, seed))
        for s in streams:
            s.synchronize()
        return outs
    for i in range(2):
        once(i)
    t0 = time.perf_counter()
    reps = 3
    for i in range(reps):
        once(10 + i)
    dt = (time.perf_counter() - t0) / reps
    del engines
    torch.cuda.empty_cache()
    return dt

for n in (1, 2):
    parts = list(cond.chunk(n))
    dt = run(parts)
    print(f"{n} concurrent sub-batches: {dt*1e3:8.1f} ms per generate -> {64*N/dt:9.0f} tok/s", flush=True)
