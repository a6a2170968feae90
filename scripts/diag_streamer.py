"""Timeline of HostStreamer legs (H2D / sweep / D2H) for one workload — diagnostic."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import simplerecon_b200 as S  # noqa: E402
from simplerecon_b200.pipeline import HostStreamer  # noqa: E402
from simplerecon_b200.synthetic import CONFIGS, make_workload_tuple, mlp_state  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = next(c for c in CONFIGS if c.name.startswith(name))
hero = w.kind == "mlp"
if hero:
    m = S.FeatureVolumeManager(w.height, w.width, w.planes, [0, 128, 128, 1], w.channels, w.views)
    m.load_state_dict({**m.state_dict(), **mlp_state(w.views, w.channels)})
else:
    m = S.CostVolumeManager(w.height, w.width, w.planes)
m = m.cuda().eval()
pin = [{k: v.pin_memory() for k, v in make_workload_tuple(w, seed_offset=i).items()} for i in range(3)]
st = HostStreamer(m, "cuda", return_mask=hero)

# instrument: wrap the three legs with events
ev = []
orig_up, orig_down = st._upload, st._download


def up(b):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st.s_in)
    r = orig_up(b)
    e1.record(st.s_in)
    ev.append(("h2d", e0, e1, time.perf_counter()))
    return r


def down(res, slot, done):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st.s_out):
        st.s_out.wait_event(done)
        e0.record(st.s_out)
    r = orig_down(res, slot, done)
    e1.record(st.s_out)
    ev.append(("d2h", e0, e1, time.perf_counter()))
    return r


st._upload, st._download = up, down
list(st.run(pin[i % 3] for i in range(4)))
torch.cuda.synchronize()
ev.clear()
base = torch.cuda.Event(enable_timing=True)
base.record()
t0 = time.perf_counter()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
list(st.run(pin[i % 3] for i in range(n)))
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"{name}: {n} steps wall {1e3 * wall:.2f} ms -> {1e3 * wall / n:.2f} ms/step")
d2h = [(base.elapsed_time(e0), e0.elapsed_time(e1)) for kind, e0, e1, th in ev if kind == "d2h"]
h2d = [(base.elapsed_time(e0), e0.elapsed_time(e1)) for kind, e0, e1, th in ev if kind == "h2d"]
if n <= 10:
    for kind, e0, e1, th in ev:
        print(f"  {kind}: start {base.elapsed_time(e0):8.2f} ms  dur {e0.elapsed_time(e1):6.2f} ms   issued at host t={1e3 * (th - t0):7.2f} ms")
else:
    for i in range(0, len(d2h) - 10, 10):
        print(f"  steps {i:3d}-{i + 10:3d}: {(d2h[i + 10][0] - d2h[i][0]) / 10:.2f} ms/step between D2H starts; "
              f"h2d dur {sum(x[1] for x in h2d[i:i + 10]) / 10:.2f} ms, d2h dur {sum(x[1] for x in d2h[i:i + 10]) / 10:.2f} ms")
