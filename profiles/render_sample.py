"""Qualitative evidence for the device-state renderer: 4096 live 3v3 envs stepped with a shoot-heavy random stream, four of
them rasterised by fr_render, written as one PNG strip (gpurun_out/r1j_render_sample.png), with the kernel's launch time."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import fortattack_b200 as fab
from importlib import import_module
rd = import_module("emergent-multiagent-strategies_b200.render")
E, A = 4096, 6
env = fab.FortAttackBatch(E, 3, 3, max_steps=100, seed=0, device="cuda:0")
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.tensor([.12] * 7 + [.16], device="cuda")
for _ in range(30):
    acts = torch.multinomial(p, A * E, replacement=True, generator=g).view(A, E).to(torch.int32)
    obs, _, _, _ = env.step(acts, auto_reset=False)
ids = [int(i) for i in torch.nonzero((acts == 7).any(0) & (obs[:, :, 0] == 0).any(0))[:4, 0]]
imgs = rd.render_batch(obs, 3, actions=acts, env_ids=ids, draw_dead=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = torch.empty(64, 700, 700, 3, dtype=torch.uint8, device="cuda")
rd.render_batch(obs, 3, actions=acts, env_ids=list(range(64)), out=out)
e0.record(); rd.render_batch(obs, 3, actions=acts, env_ids=list(range(64)), out=out); e1.record()
torch.cuda.synchronize()
from PIL import Image
strip = torch.cat(list(imgs[:, ::2, ::2]), dim=1).cpu().numpy()          # 350 x 1400
Image.fromarray(strip).save(os.path.join(ROOT, "gpurun_out", "r1j_render_sample.png"))
ms = e0.elapsed_time(e1)
print("render_sample: envs %s; 64 frames of 700 x 700 in %.3f ms = %.1f GB/s of pixels written" % (ids, ms, 64 * 700 * 700 * 3 / ms / 1e6))
