"""Per-tile timeline of the hero kernel's CTA 0 (needs a library built with -DSRCV_TC_TIMELINE:
python -m simplerecon_b200.build --out simplerecon_b200/lib/libsrcv_b200_tl.so --extra=-DSRCV_TC_TIMELINE).

    SRCV_B200_LIB=simplerecon_b200/lib/libsrcv_b200_tl.so python scripts/hero_timeline.py > gpurun_out/hero_timeline.json
"""
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import simplerecon_b200 as S  # noqa: E402
from simplerecon_b200 import _native  # noqa: E402
from simplerecon_b200.synthetic import CONFIGS, make_workload_tuple, mlp_state, to_device  # noqa: E402

w = CONFIGS[2]
t = to_device(make_workload_tuple(w, batch=4), "cuda")
m = S.FeatureVolumeManager(w.height, w.width, w.planes, [0, 128, 128, 1], w.channels, w.views)
m.load_state_dict({**m.state_dict(), **mlp_state(w.views, w.channels)})
m = m.cuda().eval()
with torch.inference_mode():
    for _ in range(3):
        out = m(**t, return_mask=True)
torch.cuda.synchronize()
lib = _native.load()
n = 48 * 16
buf = (C.c_longlong * n)()
lib.srcv_debug_read_timeline.argtypes = [C.POINTER(C.c_longlong), C.c_int32]
rc = lib.srcv_debug_read_timeline(buf, n)
ev = ["b_start", "b_block0", "b_acquired", "b_arrived", "e_mma1_seen", "e_a2_arrived", "e_mma2_seen", "e_d2free_arrived",
      "e_done", "m_a1_seen", "m_m1_issued", "m_a2_seen", "m_m2_issued", "e_a2a_arrived", "e_mma1b_seen", "m_a2b_seen"]
rows = [[buf[i * 16 + j] for j in range(16)] for i in range(48)]
t0 = rows[8][0]
tl = [{e: rows[i][j] - t0 for j, e in enumerate(ev)} for i in range(8, 40)]
per = (rows[39][9] - rows[9][9]) / 30.0
seg = lambda a, b, da=0, db=0: sum(rows[i + db][ev.index(b)] - rows[i + da][ev.index(a)] for i in range(10, 38)) / 28.0
summary = {
    "rc": rc, "clk_per_tile": per,
    "build_block0": seg("b_start", "b_block0"), "build_wait_buffer": seg("b_block0", "b_acquired"),
    "build_rest": seg("b_acquired", "b_arrived"),
    "mma_issue_after_a1": seg("m_a1_seen", "m_m1_issued"),
    "layer1_issue_to_epilogue_sees_it": seg("m_m1_issued", "e_mma1_seen"),
    "epilogue1": seg("e_mma1_seen", "e_a2_arrived"), "epilogue1_first_half": seg("e_mma1_seen", "e_a2a_arrived"),
    "epilogue1_wait_second_half": seg("e_a2a_arrived", "e_mma1b_seen"), "layer2_first_half_then_wait": seg("m_a2_seen", "m_a2b_seen"), "a2_arrive_to_mma_sees": seg("e_a2_arrived", "m_a2_seen"),
    "layer2_issue": seg("m_a2_seen", "m_m2_issued"), "layer2_issue_to_epilogue_sees_it": seg("m_m2_issued", "e_mma2_seen"),
    "epilogue2_until_buffer_free": seg("e_mma2_seen", "e_d2free_arrived"), "epilogue2_rest": seg("e_d2free_arrived", "e_done"),
    "builder_arrive_to_mma_sees_a1": seg("b_arrived", "m_a1_seen"),
    "buffer_free_to_builder_acquires(t+2)": seg("e_d2free_arrived", "b_acquired", 0, 2),
}
out = json.dumps({"summary_clk": summary, "tiles_8_to_39": tl})
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(out)
else:
    print(out)
