"""RungeKutta4 step (b200sim_step_rk4) beside the semi-implicit step and one b200sim_dynamics launch, batch 4096."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import jaxsim_b200.api as js
from jaxsim_b200 import models

dev = torch.device("cuda:0")


def timeit(f, reps=50):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for dtype in (torch.float32, torch.float64):
    for B in (4096, 16384):
        sie = js.model.JaxSimModel.build_from_model_description(models.urdf("icub_like"), time_step=1e-3)
        rk4 = js.model.JaxSimModel.build_from_model_description(models.urdf("icub_like"), time_step=1e-3,
                                                                integrator=js.model.IntegratorType.RungeKutta4)
        d = js.data.random_model_data(sie, batch_size=B, dtype=dtype, device=dev, velocity_representation=js.common.VelRepr.Inertial)
        tau = torch.zeros(B, sie.dofs(), dtype=dtype, device=dev)
        state = dict(base_position=d._base_position, base_quaternion=d._base_quaternion, joint_positions=d._joint_positions,
                     base_linear_velocity=d._base_linear_velocity, base_angular_velocity=d._base_angular_velocity,
                     joint_velocities=d._joint_velocities, contact_state=d.contact_state)
        t_sie = timeit(lambda: js.model.step(sie, d, joint_force_references=tau))
        t_dyn = timeit(lambda: js.ode.system_dynamics(sie, state, joint_torques=tau))
        t_rk4 = timeit(lambda: js.model.step(rk4, d, joint_force_references=tau))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            js.model.step(rk4, d, joint_force_references=tau)  # the stage scratch of the capture stream
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            js.model.step(rk4, d, joint_force_references=tau)
        t_graph = timeit(g.replay)
        print(f"{str(dtype):14s} B={B:6d}: semi-implicit {t_sie:7.1f} us | one system_dynamics {t_dyn:7.1f} us | "
              f"RK4 step eager {t_rk4:7.1f} us, CUDA graph {t_graph:7.1f} us ({B / t_graph:.1f} env-steps/us)")
