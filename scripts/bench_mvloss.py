"""Times MVDepthLoss forward + backward (csrc/srcv_mvloss.cu) at the training resolution against the
reference's op sequence run as PyTorch CUDA ops on the same GPU (the oracle port).  CUDA events,
warm-up, rotating inputs.  Not part of the bench.py contract (BASELINE's metric is the forward sweep).

    python scripts/bench_mvloss.py [--batch 8] [--views 7] [--height 192] [--width 256] [--steps 20]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import simplerecon_b200 as S  # noqa: E402
from oracle import mvdepth_oracle as M  # noqa: E402  (the comparison arm: the reference's op sequence)
from simplerecon_b200.synthetic import make_mvloss_batch  # noqa: E402


def timed(fn, sets, steps, warmup):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    tot = 0.0
    for it in range(warmup + steps):
        t = sets[it % len(sets)]
        p = t["depth_pred_b1hw"].clone().requires_grad_(True)
        ev[0].record()
        loss = fn(**{**t, "depth_pred_b1hw": p})
        loss.backward()
        ev[1].record()
        torch.cuda.synchronize()
        if it >= warmup:
            tot += ev[0].elapsed_time(ev[1])
    return tot / steps, float(loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--views", type=int, default=7)
    ap.add_argument("--height", type=int, default=192)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    sets = [{k: v.cuda() for k, v in make_mvloss_batch(50 + i, a.batch, a.views, a.height, a.width).items()} for i in range(3)]
    ours = S.MVDepthLoss(a.height, a.width)
    ms_ours, l_ours = timed(ours, sets, a.steps, 3)
    ms_port, l_port = timed(M.mv_depth_loss, sets, max(3, a.steps // 4), 2)
    px = a.batch * a.height * a.width
    alg = px * (8 + 4 * a.views) + px * (8 + 4 * a.views + 4)          # forward reads + backward reads / write
    print(json.dumps({"what": "MVDepthLoss forward + backward", "batch": a.batch, "views": a.views, "hw": [a.height, a.width],
                      "ours_ms": round(ms_ours, 4), "reference_ops_on_this_gpu_ms": round(ms_port, 3),
                      "speedup": round(ms_port / ms_ours, 1), "loss_ours": l_ours, "loss_port": l_port,
                      "algorithmic_bytes": alg, "GBps_at_ours_ms": round(alg / ms_ours / 1e6, 1),
                      "note": "launch-latency-sized: 3 launches (+ the dense fp32 input copies of the Python mirror)"}))


if __name__ == "__main__":
    main()
