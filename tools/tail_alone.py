"""The layer-tail launches of one decode position timed alone (attention skipped), per model and row count:
    python tools/tail_alone.py GPT-XL 2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg import _lib as L
from vlg.gpt import GPT_models

name, rows = sys.argv[1], int(sys.argv[2])
torch.manual_seed(0)
m = GPT_models[name](block_size=576, cls_token_num=1, model_type="c2i").eval().to("cuda", torch.bfloat16)
eng = m.engine(rows, 577)
eng.begin(rows)
lib = L.lib()
st = torch.cuda.Stream()
cfg = m.config
hid = m.layers[0].feed_forward.w1.weight.shape[0]
with torch.cuda.stream(st):
    L.check(lib.vlg_debug_trunk_without_attention(eng.ctx, 100, L.stream()))
    st.synchronize()
    graph = torch.cuda.CUDAGraph()
    reps = 8
    with torch.cuda.graph(graph, stream=st):
        for i in range(reps):
            L.check(lib.vlg_debug_trunk_without_attention(eng.ctx, 100 + i, L.stream()))
    graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.synchronize()
    e0.record(st)
    graph.replay()
    e1.record(st)
    st.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * (cfg.n_layer + 1))
w_mb = (4 * cfg.dim * cfg.dim + 3 * cfg.dim * hid) * 2 / 1e6
print(f"{name} {rows} rows: {us:.2f} us per layer-tail launch ({w_mb:.1f} MB of weights = {w_mb / us * 1e3 / 1e3:.2f} TB/s; HBM floor {w_mb / 6544.3 * 1e3:.2f} us)")
