"""CPU: the TSDF integration kernel (csrc/srcv_tsdf.cu) compiled for the host (tests/emu) and the
Python mirror of the reference's TSDF / TSDFFuser classes, bit for bit against the oracle."""
import contextlib
import types

import pytest
import torch

from oracle import tsdf_oracle as T
from simplerecon_b200 import _native, tsdf as tsdf_mod
from simplerecon_b200.synthetic import make_tsdf_case
from tests import emu


@pytest.fixture()
def emulated(monkeypatch):
    lib = emu.load_or_skip()
    monkeypatch.setattr(_native, "_lib", lib)
    monkeypatch.setattr(tsdf_mod, "_require_cuda", lambda t: None)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: types.SimpleNamespace(cuda_stream=0))
    real_empty = torch.empty

    def aligned_empty(*size, **kw):
        if kw.get("dtype") is torch.uint8 and len(size) == 1 and isinstance(size[0], int):
            buf = real_empty(size[0] + 256, **kw)
            off = (-buf.data_ptr()) % 256
            return buf[off:off + size[0]]
        return real_empty(*size, **kw)

    monkeypatch.setattr(torch, "empty", aligned_empty)
    return lib


@pytest.mark.parametrize("seed,frames,voxel,masked", [(5, 2, 0.07, False), (6, 3, 0.09, True)])
def test_integrate_matches_oracle_bitwise(emulated, seed, frames, voxel, masked):
    c = make_tsdf_case(seed=seed, frames=frames, voxel_size=voxel, height=48, width=64, masked=masked)
    vol = tsdf_mod.TSDF.from_bounds(c["bounds"], voxel, device="cpu")
    fuser = tsdf_mod.TSDFFuser(vol, max_depth=c["max_depth"])
    tv, tw, origin = T.new_volume(c["bounds"], voxel)
    assert tuple(tv.shape) == tuple(vol.tsdf_values.shape) and torch.equal(origin, vol.origin)
    assert torch.equal(vol.voxel_coords, T.voxel_coords(origin, tuple(tv.shape), voxel))
    for _ in range(2):
        fuser.integrate_depth(c["depth"], c["cam_T_world"], c["K"], c["mask"])
        T.integrate(tv, tw, origin, voxel, c["depth"], c["cam_T_world"], c["K"], c["mask"],
                    min_depth=fuser.min_depth, max_depth=c["max_depth"])
        assert int((tw > 0).sum()) > 1000
        assert torch.equal(vol.tsdf_weights, tw), (vol.tsdf_weights.float() - tw.float()).abs().max()
        assert torch.equal(vol.tsdf_values, tv), (vol.tsdf_values.float() - tv.float()).abs().max()


def test_scalar_path_and_argument_checks(emulated):
    """Z not a multiple of 8 takes the one-voxel-per-thread path; bad arguments are refused."""
    c = make_tsdf_case(seed=7, frames=1, voxel_size=0.1, height=24, width=32)
    tv, tw, origin = T.new_volume(c["bounds"], 0.1)
    tv, tw = tv[:, :, :29].contiguous(), tw[:, :, :29].contiguous()
    vol = tsdf_mod.TSDF(tv.clone(), tw.clone(), 0.1, origin)
    tsdf_mod.TSDFFuser(vol, max_depth=3.0).integrate_depth(c["depth"], c["cam_T_world"], c["K"])
    T.integrate(tv, tw, origin, 0.1, c["depth"], c["cam_T_world"], c["K"], max_depth=3.0)
    assert int((tw > 0).sum()) > 100
    assert torch.equal(vol.tsdf_weights, tw) and torch.equal(vol.tsdf_values, tv)
    with pytest.raises(_native.SrcvError):
        tsdf_mod.TSDFFuser(vol, min_depth=3.0, max_depth=3.0).integrate_depth(c["depth"], c["cam_T_world"], c["K"])
