"""Profiling driver (run under ncu): BASELINE configs[4] scoring, ML-20M shape, 10 000 eval users."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "revisit-bpr_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from rbpr import synth  # noqa: E402
from rbpr.engine import Engine  # noqa: E402

dev = torch.device("cuda:0")
inter = bench.load_interactions("ml-20m", 1.0)
g = torch.Generator(device=dev).manual_seed(13)
ue = torch.randn(inter.num_users, 128, generator=g, device=dev) * 0.1
ie = torch.randn(inter.num_items, 128, generator=g, device=dev) * 0.1
ue[0] = 0
ie[0] = 0
eng = Engine(ue, ie)
users, seen, held = synth.split_heldout(inter, 10_000)
t = lambda a, dt: torch.from_numpy(a).to(dev, dt)  # noqa: E731
args = (t(users, torch.int64), (t(seen[0], torch.int64), t(seen[1], torch.int32)), (t(held[0], torch.int64), t(held[1], torch.int32)))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = eng.score_metrics(*args, [20, 100], want=("ndcg", "recall"))
torch.cuda.synchronize()
print("tensor passes, overflow users:", eng.score_path_counts(), "ndcg@100", float(out["ndcg"][:, 1].mean()))
