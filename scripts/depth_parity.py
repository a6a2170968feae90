"""Depth-map parity through the reference's own U-Net (BASELINE target: depth Abs-Diff <= 1e-4).

The reference's `CVEncoder` + `DepthDecoderPP` (reference modules/networks.py:20-127) turn a
cost volume + image-prior features into the depth map.  They live only in /root/reference
(build container, CPU); our kernels run only on the GPU box.  So the check has two halves:

    # on the GPU box (through gpurun): our cost volumes for seeded 640x480 tuples
    python scripts/depth_parity.py dump        ->  gpurun_out/depth_parity_ours.npz

    # in the build container: reference cost volumes (CPU, reference classes) for the SAME
    # tuples, then the reference decoder on both, Abs-Diff of depth_pred_s0_b1hw
    python scripts/depth_parity.py eval        ->  profiles/depth_parity_r01.json

Seeded random weights everywhere (no checkpoint exists in this environment);
image-prior pyramid = seeded random tensors with EfficientNetV2-S channel counts
[24,48,64,160,256] at strides 2..32 (reference depth_model.py:108-116).  Abs-Diff is
`mean|a-b|` (reference utils/metrics_utils.py:38).
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from simplerecon_b200.synthetic import make_tuple, mlp_state  # noqa: E402

H, W, D, K, C = 120, 160, 64, 7, 16
SEEDS = {"dot": 2024, "hero": 2025}
DUMP = ROOT / "gpurun_out" / "depth_parity_ours.npz"


def dump():
    import simplerecon_b200 as S
    from simplerecon_b200 import _native
    from simplerecon_b200.synthetic import to_device
    out = {}
    with torch.inference_mode():
        for kind, seed in SEEDS.items():
            t = to_device(make_tuple(1, K, H, W, seed=seed, smooth=True), "cuda")
            if kind == "dot":
                m = S.CostVolumeManager(H, W, D).cuda()
            else:
                m = S.FeatureVolumeManager(H, W, D, [0, 128, 128, 1], C, K)
                m.load_state_dict({**m.state_dict(), **mlp_state(K, C, seed=0)})
                m = m.cuda()
            cost, lowest, planes, mask = m(**t, return_mask=True)
            out[f"{kind}_cost"] = cost.cpu().numpy()
            out[f"{kind}_lowest"] = lowest.cpu().numpy()
            out[f"{kind}_variant"] = np.array(_native.last_variant())
    DUMP.parent.mkdir(exist_ok=True)
    np.savez_compressed(DUMP, **out)
    print("wrote", DUMP, {k: getattr(v, "shape", None) for k, v in out.items()})


def evaluate():
    from oracle.ref_import import load_reference
    R = load_reference()
    z = np.load(DUMP)
    torch.manual_seed(0)
    enc_ch = [24, 48, 64, 160, 256]                      # EfficientNetV2-S feature_info.channels()
    cv_net = R.CVEncoder(num_ch_cv=D, num_ch_enc=enc_ch[1:], num_ch_outs=[64, 128, 256, 384]).eval()
    dec = R.DepthDecoderPP(enc_ch[:1] + cv_net.num_ch_enc).eval()
    g = torch.Generator().manual_seed(7)
    img_feats = [torch.randn(1, c, 4 * H // (2 ** (i + 1)), 4 * W // (2 ** (i + 1)), generator=g)
                 for i, c in enumerate(enc_ch)]

    def depth_from(cv):
        with torch.no_grad():
            feats = img_feats[:1] + cv_net(cv, img_feats[1:])
            outs = dec(feats)
        return {k.replace("log_", ""): torch.exp(v.float()) for k, v in outs.items()}

    report = {}
    for kind, seed in SEEDS.items():
        t = make_tuple(1, K, H, W, seed=seed, smooth=True)
        if kind == "dot":
            ref = R.CostVolumeManager(H, W, num_depth_bins=D)
        else:
            ref = R.FeatureVolumeManager(H, W, num_depth_bins=D, mlp_channels=[0, 128, 128, 1],
                                         matching_dim_size=C, num_source_views=K)
            ref.load_state_dict({**ref.state_dict(), **mlp_state(K, C, seed=0)})
        with torch.no_grad():
            rc, rl, _, _ = ref(**t, return_mask=True)
            rc64, _, _, _ = ref.double()(**{k: v.double() for k, v in t.items()}, return_mask=True)
        ours = torch.from_numpy(z[f"{kind}_cost"])
        d_ref, d_ours, d_64 = depth_from(rc), depth_from(ours), depth_from(rc64.float())
        key = "depth_pred_s0_b1hw"
        report[kind] = {
            "kernel_variant": str(z[f"{kind}_variant"]),
            "cost_max_abs": float(rc.abs().max()),
            "cost_err_ours_vs_ref32": float((ours - rc).abs().max()),
            "cost_err_ours_vs_ref64": float((ours.double() - rc64).abs().max()),
            "cost_err_ref32_vs_ref64": float((rc.double() - rc64).abs().max()),
            # argmax plane index (plane VALUES differ by an ulp between the CPU's and the GPU's exp/log)
            "argmax_plane_mismatch_px": int((ours.argmax(1) != rc.argmax(1)).sum()),
            "lowest_cost_max_rel_diff": float(((torch.from_numpy(z[f"{kind}_lowest"]) - rl).abs() / rl).max()),
            "depth_shape": list(d_ref[key].shape),
            "depth_mean": float(d_ref[key].mean()),
            "depth_abs_diff_ours_vs_ref": float((d_ours[key] - d_ref[key]).abs().mean()),
            "depth_max_diff_ours_vs_ref": float((d_ours[key] - d_ref[key]).abs().max()),
            # the reference's own fp32-vs-fp64 cost-volume noise pushed through the same decoder
            "depth_abs_diff_ref32_vs_ref64cv": float((d_ref[key] - d_64[key]).abs().mean()),
            "target_abs_diff": 1e-4,
        }
        report[kind]["pass"] = report[kind]["depth_abs_diff_ours_vs_ref"] <= 1e-4
    out = ROOT / "profiles" / "depth_parity_r01.json"
    out.write_text(json.dumps(report, indent=1))
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    {"dump": dump, "eval": evaluate}[sys.argv[1]]()
