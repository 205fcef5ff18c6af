import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import sift_project_b200 as S
from oracle import oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
na = nb = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
a, b = O.synth_descriptors(na, seed=11), O.synth_descriptors(nb, seed=12)
ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
fa, fb = ta.double(), tb.double()
D = (fa * fa).sum(1)[:, None] + (fb * fb).sum(1)[None, :] - 2 * fa @ fb.T
D = D.round().long()
# stable top-2: sort by (d, j)
key = D * 65536 + torch.arange(nb, device="cuda")[None, :]
k2, _ = torch.topk(key, 2, dim=1, largest=False)
wi = (k2[:, 0] % 65536).cpu().numpy(); w1 = (k2[:, 0] // 65536).cpu().numpy(); w2 = (k2[:, 1] // 65536).cpu().numpy()
ctx = S.SiftContext(64, 64)
for path in ("tc", "simt"):
    os.environ["SIFT_B200_MATCH"] = path
    idx = torch.empty(na, dtype=torch.int32, device="cuda"); d1 = torch.empty_like(idx); d2 = torch.empty_like(idx)
    torch.cuda.synchronize()
    ctx.match_enqueue(ta, na, tb, nb, idx, d1, d2); ctx.sync()
    gi, g1, g2 = idx.cpu().numpy(), d1.cpu().numpy(), d2.cpu().numpy()
    bad = np.nonzero((gi != wi) | (g1 != w1) | (g2 != w2))[0]
    print(path, "mismatches", len(bad))
    for r in bad[:10]:
        print("  row", r, "got", gi[r], g1[r], g2[r], "want", wi[r], w1[r], w2[r])
    # timing
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3): ctx.match_enqueue(ta, na, tb, nb, idx, d1, d2)
    ctx.sync()
    ev0.record(st)
    for _ in range(10): ctx.match_enqueue(ta, na, tb, nb, idx, d1, d2)
    ev1.record(st); ctx.sync()
    ms = ev0.elapsed_time(ev1) / 10
    print(path, f"{ms*1e3:.1f} us  {2*na*nb*128/ms/1e9:.1f} TFLOP/s-equivalent")
