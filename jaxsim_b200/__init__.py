"""jaxsim_b200 -- B200-native batched rigid-body simulation step.

A drop-in for ONE hot path of ami-iit/jaxsim: ``jaxsim.api.model.step`` vmapped over
environments (and the functions under it), executed by hand-written sm_100a CUDA kernels
behind the C ABI of ``include/b200sim.h``.  Usage mirrors the reference (README.md:39-84):

    import jaxsim_b200 as jaxsim
    import jaxsim_b200.api as js
    model = js.model.JaxSimModel.build_from_model_description(urdf, time_step=1e-3)
    data = js.data.random_model_data(model, batch_size=4096, dtype=torch.float32)
    data = js.model.step(model, data, joint_force_references=tau)
"""

from . import api, models, rbda, terrain  # noqa: F401
from .api.common import VelRepr  # noqa: F401

__version__ = "0.1.0"
