"""Times one training step (fused forward + backward kernel) of the managers on the GPU:
CUDA events, warm-up, rotating inputs.  Not part of the bench.py contract (BASELINE's metric
is the forward sweep); this is the measurement script for the §8f-1 backward kernels.

    python scripts/bench_backward.py [--workload cfg1|cfg2] [--batch B] [--steps N]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import simplerecon_b200 as S  # noqa: E402
from simplerecon_b200 import _native  # noqa: E402
from simplerecon_b200.synthetic import make_tuple, mlp_state, to_device  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    K, C, H, W, D = 7, 16, 120, 160, 64
    B = a.batch or (4 if a.workload == "cfg1" else 8)
    if a.workload == "cfg1":
        m = S.CostVolumeManager(H, W, num_depth_bins=D).cuda()
    else:
        m = S.FeatureVolumeManager(H, W, num_depth_bins=D, mlp_channels=[0, 128, 128, 1], matching_dim_size=C,
                                   num_source_views=K)
        m.load_state_dict({**m.state_dict(), **mlp_state(views=K, channels=C, seed=0)})
        m = m.cuda().train()
    sets = [to_device(make_tuple(B, K, H, W, channels=C, seed=100 + i), "cuda") for i in range(3)]
    gcost = torch.randn(B, D, H, W, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd_ms = bwd_ms = 0.0
    for it in range(a.warmup + a.steps):
        t = dict(sets[it % len(sets)])
        t["cur_feats"] = t["cur_feats"].detach().requires_grad_(True)
        t["src_feats"] = t["src_feats"].detach().requires_grad_(True)
        m.zero_grad(set_to_none=True)
        ev[0].record()
        cost, *_ = m(**t)
        ev[1].record()
        cost.backward(gcost)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= a.warmup:
            fwd_ms += ev[0].elapsed_time(ev[1])
            bwd_ms += ev[1].elapsed_time(ev[2])
    rows = B * D * H * W
    flop_bwd = rows * (258e3 if a.workload == "cfg2" else 2 * 2 * K * C * 5)
    print(json.dumps({"workload": a.workload, "batch": B, "steps": a.steps,
                      "forward_ms": round(fwd_ms / a.steps, 4), "backward_ms": round(bwd_ms / a.steps, 4),
                      "backward_variant": _native.last_variant(),
                      "backward_tflops_fp32": round(flop_bwd / (bwd_ms / a.steps * 1e-3) / 1e12, 2)}))


if __name__ == "__main__":
    main()
