import os, sys, torch
sys.path.insert(0, "/root/repo")
from tec_mollm_b200 import SpatioTemporalEmbedding
dev = torch.device("cuda", 0)
B, L, N = 128, 48, 2911
emb = SpatioTemporalEmbedding(16, num_nodes=N).to(dev)
x = torch.randn(B, L, N, 6, device=dev)
tf = torch.stack([torch.randint(0, 12, (B, L)), torch.randint(0, 366, (B, L)), torch.randint(0, 13, (B, L)), torch.randint(0, 4, (B, L))], -1).float().to(dev)
for npb in (128, 256, 384, 512, 1024, 2911):
    os.environ["TECGAT_EMBED_NPB"] = str(npb)
    with torch.no_grad():
        for _ in range(3): emb(x, tf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): emb(x, tf)
        e1.record(); torch.cuda.synchronize()
    print(npb, round(e0.elapsed_time(e1) / 20, 4), "ms", flush=True)
