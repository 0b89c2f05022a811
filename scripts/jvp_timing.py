"""Where the time of BASELINE config 5 goes: step_jvp at batch 4096 with and without a mass direction, the same kernel on
K x 4096 replicated environments (K state directions in one launch), and the float64 primal step beside it."""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import jaxsim_b200.api as js
from jaxsim_b200 import models

dev = torch.device("cuda:0")
model = js.model.JaxSimModel.build_from_model_description(models.urdf("icub_like"), time_step=1e-3)
B = 4096
d = js.data.random_model_data(model, batch_size=B, dtype=torch.float64, device=dev, velocity_representation=js.common.VelRepr.Inertial)
n, nL = model.dofs(), model.number_of_links()


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps


tq = torch.randn(B, n, dtype=torch.float64, device=dev)
tm = torch.randn(nL, dtype=torch.float64)
print("primal f64 step            : %.3f ms (wall %.3f)" % timeit(lambda: js.model.step(model, d)))
print("jvp, joint direction       : %.3f ms (wall %.3f)" % timeit(lambda: js.model.step_jvp(model, d, {"joint_positions": tq})))
print("jvp, joint + mass direction: %.3f ms (wall %.3f)" % timeit(lambda: js.model.step_jvp(model, d, {"joint_positions": tq, "link_masses": tm})))
for K in (4, 8, 30):
    from jaxsim_b200.api.data import _map_leaves
    dK = _map_leaves(d, lambda t: t.repeat((K,) + (1,) * (t.dim() - 1)))
    tK = torch.randn(K * B, n, dtype=torch.float64, device=dev)
    ev, wall = timeit(lambda: js.model.step_jvp(model, dK, {"joint_positions": tK}), reps=5)
    print("jvp, %2d directions as one batch of %6d: %.3f ms (wall %.3f) = %.3f ms per direction" % (K, K * B, ev, wall, ev / K))
print(model.launch_geometry(B, torch.float64, dev))
